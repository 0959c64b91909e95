#!/usr/bin/env python
"""per-family CUDA-event times of bader_calc('ongrid') on an n^3 bench grid"""
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
for it in range(3):
    e.clear_labels(LABELS_BADER); e.bader_calc('ongrid', dist, T)
e.profile(True); e.profile_reset()
for it in range(3):
    e.clear_labels(LABELS_BADER); mx = e.bader_calc('ongrid', dist, T)
print(os.environ.get('BDR_RESOLVE_MODE'), len(mx), {k: round(v[0] / v[1], 3) for k, v in e.profile_get().items()})
