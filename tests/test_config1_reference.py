"""BASELINE configs 1 and 2 at full size against the REAL reference: tests/golden_c1/*.npz hold
what pybader's numba kernels returned for the 96^3 three-atom cell (make_c1_golden.py).
The density is regenerated and must hash to the stored value, else these tests skip."""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def c1():
    from pybader_b200 import geometry as geo, synth
    g = np.load(os.path.join(ROOT, 'tests', 'golden_c1', 'c1_96.npz'))
    c = synth.case_c1(96)
    rho, _ = synth.make(c)
    if hashlib.sha256(rho.tobytes()).digest() != g['rho_sha256'].tobytes():
        pytest.skip("this CPU's exp() does not reproduce the stored density bit for bit")
    return dict(g=g, rho=rho, dist=geo.distance_matrix(c['lattice'], rho.shape),
                T=geo.T_grad(c['lattice'], rho.shape), dV=geo.voxel_volume(c['lattice'], rho.shape))


def test_oracle_equals_reference_at_full_size(c1):
    """the C oracle reproduces the reference bit for bit at 96^3: ongrid, the scan-order
    dependent raw neargrid labels, the refined labels, the sums"""
    from oracle import pyoracle as orc
    g, rho = c1['g'], c1['rho']
    mx, vol = orc.bader_calc('ongrid', rho, np.zeros(rho.shape, np.int32), c1['dist'], c1['T'])
    np.testing.assert_array_equal(mx, g['ongrid_maxima'])
    assert vol.dtype == g['ongrid_labels'].dtype
    np.testing.assert_array_equal(vol, g['ongrid_labels'])
    mx, vol = orc.bader_calc('neargrid', rho, np.zeros(rho.shape, np.int32), c1['dist'], c1['T'])
    np.testing.assert_array_equal(mx, g['neargrid_maxima'])
    np.testing.assert_array_equal(vol, g['neargrid_raw_labels'])
    orc.refine('neargrid', ('changed', 2), rho, vol, c1['dist'], c1['T'])
    np.testing.assert_array_equal(vol, g['neargrid_refined_labels'])
    n = mx.shape[0]
    q, v = np.zeros(n), np.zeros(n)
    orc.charge_sum(q, v, c1['dV'], rho, vol)
    np.testing.assert_array_equal(q, g['charge'])
    np.testing.assert_array_equal(v, g['volume'])


@pytest.mark.gpu
def test_cuda_path_equals_reference_at_full_size(c1):
    """the CUDA path against the reference's own output: ongrid bit-exact; neargrid +
    refine ('changed', 2) >= 99.9 % of the labels, charges and volumes to 1e-6"""
    from pybader_b200 import thread_handlers as th, utils as ut
    g, rho = c1['g'], c1['rho']
    mx, vol = th.bader_calc('ongrid', rho, np.zeros(rho.shape, np.int32), c1['dist'], c1['T'], 1)
    np.testing.assert_array_equal(mx, g['ongrid_maxima'])
    assert vol.dtype == g['ongrid_labels'].dtype
    np.testing.assert_array_equal(vol, g['ongrid_labels'])
    mx, vol = th.bader_calc('neargrid', rho, np.zeros(rho.shape, np.int32), c1['dist'], c1['T'], 1)
    th.refine('neargrid', ('changed', 2), rho, vol, c1['dist'], c1['T'], 1)
    np.testing.assert_array_equal(mx, g['neargrid_maxima'])      # same maxima, same numbering
    agree = np.mean(vol == g['neargrid_refined_labels'])
    assert agree >= 0.999, agree
    n = mx.shape[0]
    q, v = np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, c1['dV'], rho, vol)
    np.testing.assert_allclose(q, g['charge'], rtol=1e-6)
    np.testing.assert_allclose(v, g['volume'], rtol=1e-6)
    print(f"config 1 vs the reference itself: {int((vol != g['neargrid_refined_labels']).sum())} "
          f"of {vol.size} voxels differ")


# ---------------------------------------------------------------- config 2 ----
@pytest.fixture(scope='module')
def c2():
    from pybader_b200 import geometry as geo, synth
    g = np.load(os.path.join(ROOT, 'tests', 'golden_c1', 'c2_256.npz'))
    c = synth.case_rocksalt(256, cells=4, offset=0.13)
    tx, ty, tz = synth.separable_tables(c)
    rho = np.ascontiguousarray(np.einsum('ai,aj,ak->ijk', tx, ty, tz, optimize=True))
    if hashlib.sha256(rho.tobytes()).digest() != g['rho_sha256'].tobytes():
        pytest.skip("this CPU does not reproduce the stored density bit for bit")
    return dict(g=g, rho=rho, dist=geo.distance_matrix(c['lattice'], rho.shape),
                T=geo.T_grad(c['lattice'], rho.shape))


def test_oracle_equals_reference_config2(c2):
    """BASELINE config 2 (256^3 rocksalt-like cell, 64 maxima, method=ongrid): the oracle
    against the real reference's maxima list and labels, bit for bit"""
    from oracle import pyoracle as orc
    g, rho = c2['g'], c2['rho']
    mx, vol = orc.bader_calc('ongrid', rho, np.zeros(rho.shape, np.int32), c2['dist'], c2['T'])
    np.testing.assert_array_equal(mx, g['ongrid_maxima'])
    assert vol.dtype == g['ongrid_labels'].dtype
    np.testing.assert_array_equal(vol, g['ongrid_labels'])


@pytest.mark.gpu
def test_cuda_path_equals_reference_config2(c2):
    """the CUDA ongrid path against the real reference's output at 256^3: bit-exact"""
    from pybader_b200 import thread_handlers as th
    g, rho = c2['g'], c2['rho']
    mx, vol = th.bader_calc('ongrid', rho, np.zeros(rho.shape, np.int32), c2['dist'], c2['T'], 1)
    np.testing.assert_array_equal(mx, g['ongrid_maxima'])
    assert vol.dtype == g['ongrid_labels'].dtype
    np.testing.assert_array_equal(vol, g['ongrid_labels'])
