#!/usr/bin/env python
"""per-family CUDA-event times of one neargrid+refine step on an n^3 bench grid"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from pybader_b200 import geometry as geo, synth
from pybader_b200.engine import Engine, LABELS_BADER
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
shape = (n, n, n)
case, _ = B.workload_case(shape)
dist = geo.distance_matrix(case['lattice'], shape); T = geo.T_grad(case['lattice'], shape)
e = Engine(shape); e.synth_separable(0, *synth.separable_tables(case))
def step():
    e.clear_labels(LABELS_BADER); mx = e.bader_calc('neargrid', dist, T)
    return e.refine(LABELS_BADER, 'changed', 2, dist, T)
for it in range(2): step()
e.profile(True); e.profile_reset()
e.timer_start()
for it in range(2): h = step()
ms = e.timer_stop() / 2
print(os.environ.get('TAG', ''), 'step %.2f ms' % ms, {k: round(v[0] / 2, 2) for k, v in e.profile_get().items()}, h)
