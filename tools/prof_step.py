#!/usr/bin/env python
"""One neargrid + refine step on a cubic grid, for ncu captures:
    ncu --set full -k regex:... python tools/prof_step.py 512"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from pybader_b200 import geometry as geo, synth  # noqa: E402
from pybader_b200.engine import Engine, LABELS_BADER  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape = (n, n, n)
case, _ = B.workload_case(shape)
dist = geo.distance_matrix(case['lattice'], shape)
T = geo.T_grad(case['lattice'], shape)
e = Engine(shape)
e.synth_separable(0, *synth.separable_tables(case))
for _ in range(steps):
    e.clear_labels(LABELS_BADER)
    mx = e.bader_calc('neargrid', dist, T)
    hist = e.refine(LABELS_BADER, 'changed', 2, dist, T)
print(len(mx), hist)
