#!/usr/bin/env python
"""Text -> grid throughput of the CHGCAR reader next to the
reference's conversion (numpy turning the split tokens into float64, io/vasp.py:94-103)
on a bounded sample of the same file.  One JSON line.

    python tools/io_bench.py [N]      # N^3 grid, default 256
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybader_b200.io import vasp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(5)
N = n ** 3
vals = rng.lognormal(0, 2, N)
path = os.path.join(tempfile.gettempdir(), f'CHGCAR_bench_{n}')
t0 = time.perf_counter()
with open(path, 'w') as f:
    f.write("bench\n 1.0\n 10.0 0.0 0.0\n 0.0 10.0 0.0\n 0.0 0.0 10.0\n H\n 1\nDirect\n 0.0 0.0 0.0\n\n")
    f.write(f" {n} {n} {n}\n")
    full = N // 5 * 5
    np.savetxt(f, vals[:full].reshape(-1, 5), fmt='%18.11E', delimiter=' ')
    if full < N:
        np.savetxt(f, vals[full:].reshape(1, -1), fmt='%18.11E', delimiter=' ')
t_write = time.perf_counter() - t0
size = os.path.getsize(path)
with contextlib.redirect_stdout(io.StringIO()):
    vasp.read(path)                         # warm: CUDA context, page cache
    t0 = time.perf_counter()
    d, *_ = vasp.read(path)
    t_gpu = time.perf_counter() - t0
# the reference's conversion on a bounded sample of the same text
m = min(N, 2_000_000)
with open(path, 'rb') as f:
    for _ in range(11):
        f.readline()
    sample = f.read(m * 19)
t0 = time.perf_counter()
ref = np.zeros(m)
ref[:] = sample.decode().strip().split()[:m]
t_ref = time.perf_counter() - t0
x = np.swapaxes(vals.reshape(n, n, n), 0, -1) / 1000.0
ok = bool(np.allclose(d['charge'], x, rtol=1e-10))
os.remove(path)
print(json.dumps({
    "what": f"CHGCAR reader, {n}^3 grid, {size / 1e6:.0f} MB of text",
    "gpu_reader_s": t_gpu, "gpu_values_per_s": N / t_gpu, "gpu_text_GBps": size / t_gpu / 1e9,
    "reference_conversion_values_per_s": m / t_ref, "reference_sample_values": m,
    "speedup_on_conversion": (N / t_gpu) / (m / t_ref), "values_match": ok,
    "note": "gpu_reader_s is the whole read(): file -> host bytes -> H2D -> tokenise/convert/transpose -> D2H numpy"}))
