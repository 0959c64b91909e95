"""One rank of a sharded run on ONE GPU (world size 1: the slab ring closes on itself), to look
at the per-rank kernels of a given slab shape without paying for N GPUs.
    python tools/slab_probe.py NX NY NZ [halo] [steps]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from pybader_b200 import geometry as geo, synth  # noqa: E402
from pybader_b200.engine import FAMILIES  # noqa: E402
from pybader_b200.sharded import Comm, ShardedBader, SlabBackend  # noqa: E402

shape = tuple(int(a) for a in sys.argv[1:4])
halo = int(sys.argv[4]) if len(sys.argv) > 4 else 4
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
os.environ.setdefault('MASTER_PORT', '29533')
torch.cuda.set_device(0)
dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', 0))
case, cells = B.workload_case(shape)
dm, T = geo.distance_matrix(case['lattice'], shape), geo.T_grad(case['lattice'], shape)
sb = ShardedBader(shape, Comm(), lambda ws, h: SlabBackend(ws, h, device=0), halo=halo)
tx, ty, tz = synth.separable_tables(case)
sb.backend.synth_separable(0, np.ascontiguousarray(tx[:, sb.window_x]), ty, tz)
be = sb.backend


def step():
    be.clear_labels()
    sb.neargrid(dm, T)
    return sb.refine(dm, T, 2, mode='changed')


for _ in range(2):
    step()
be.check(be.lib.bdr_profile_enable(be.h, 1))
be.check(be.lib.bdr_profile_reset(be.h))
torch.cuda.synchronize()
be.timer_start()
for _ in range(steps):
    hist = step()
ms = be.timer_stop() / steps
prof = {}
for i, name in enumerate(FAMILIES):
    pm, pn = ctypes.c_double(0), ctypes.c_int64(0)
    be.check(be.lib.bdr_profile_get(be.h, i, ctypes.byref(pm), ctypes.byref(pn)))
    if pn.value:
        prof[name] = (round(pm.value / steps, 3), pn.value / steps)
print(json.dumps({"shape": shape, "halo": halo, "ms_per_step": ms, "kernels": prof, "hist": hist,
                  "rounds": sb.neargrid_history}))
dist.destroy_process_group()
