#!/usr/bin/env python
"""BASELINE config 2 at full size through the REAL reference: rocksalt-like 64-atom cell
256^3, method=ongrid, threads=1.  Density regenerated from the separable tables of
pybader_b200.synth.case_rocksalt(256) and pinned by a checksum; stored: maxima, labels (int8).

    python tests/golden_c1/make_c2_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden import import_reference, quiet  # noqa: E402

from pybader_b200 import geometry as geo, synth  # noqa: E402

ref = import_reference()
c = synth.case_rocksalt(256, cells=4, offset=0.13)
tx, ty, tz = synth.separable_tables(c)
rho = np.ascontiguousarray(np.einsum('ai,aj,ak->ijk', tx, ty, tz, optimize=True))
dist = geo.distance_matrix(c['lattice'], rho.shape)
T = geo.T_grad(c['lattice'], rho.shape)
with quiet():
    mx, vol = ref['th'].bader_calc('ongrid', rho, np.zeros(rho.shape, dtype=np.int32), dist, T, 1)
out = dict(rho_sha256=np.frombuffer(hashlib.sha256(rho.tobytes()).digest(), dtype=np.uint8),
           ongrid_maxima=mx, ongrid_labels=vol)
np.savez_compressed(os.path.join(HERE, 'c2_256.npz'), **out)
print(mx.shape, vol.dtype, 'file size', os.path.getsize(os.path.join(HERE, 'c2_256.npz')))
