"""GPU parity tests: the CUDA path, called through the reference-shaped entry
points and the C ABI, against (a) the committed golden fixtures produced by the
real reference and (b) the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star):
  * ongrid pointers/labels/maxima, edge classification, one trace iteration,
    edge_check and the refine drivers started from identical labels: BIT-EXACT;
  * neargrid + refinement labels: >= 99.9 % identical, every disagreement on a
    Bader-surface voxel; per-volume / per-atom charge and volume within 1e-6
    relative.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6          # charges / volumes (north_star)
LABEL_AGREE = 0.999     # neargrid / refined labels (north_star)


@pytest.fixture(scope='module')
def th():
    from pybader_b200 import build
    build.build()
    from pybader_b200 import thread_handlers
    return thread_handlers


@pytest.fixture(scope='module')
def ut():
    from pybader_b200 import utils
    return utils


@pytest.fixture(scope='module')
def orc():
    from oracle import pyoracle
    return pyoracle


def fresh_volumes(ut, g):
    shape = g['rho'].shape
    v = np.zeros(shape, dtype=ut.dtype_calc(-int(np.prod(shape))))
    if g['vacuum_tol'] is not None:
        v, q, vol = ut.vacuum_assign(g['rho'], v, g['vacuum_tol'], g['rho'], float(g['voxel_volume']))
        assert q == pytest.approx(float(g['vacuum_charge']), rel=1e-12)
        assert vol == pytest.approx(float(g['vacuum_volume']), rel=1e-9)
        assert int((v == -1).sum()) == round(float(g['vacuum_volume']) / float(g['voxel_volume']))
    return v


def canonical(labels, maxima):
    """relabel by the linear voxel index of each volume's maximum"""
    shape = labels.shape
    key = (maxima[:, 0] * shape[1] + maxima[:, 1]) * shape[2] + maxima[:, 2]
    out = np.full(labels.shape, -1, dtype=np.int64)
    sel = labels >= 0
    out[sel] = key[labels[sel]]
    return out


# ---------------------------------------------------------------- golden ----
def test_ongrid_golden_bit_exact(th, ut, golden):
    g = golden
    mx, vol = th.bader_calc('ongrid', g['rho'], fresh_volumes(ut, g), g['dist_mat'], g['T_grad'], 1)
    assert vol.dtype == g['ongrid_volumes'].dtype
    np.testing.assert_array_equal(mx, g['ongrid_maxima'])
    np.testing.assert_array_equal(vol, g['ongrid_volumes'])


def test_edge_find_and_one_trace_golden_bit_exact(golden):
    from pybader_b200.engine import Engine, LABELS_BADER
    g = golden
    e = Engine(g['rho'].shape)
    e.upload_density(0, g['rho'])
    e.upload_labels(LABELS_BADER, g['ongrid_volumes'])
    assert e.edge_find(LABELS_BADER) == int(g['ongrid_edges'])
    np.testing.assert_array_equal(e.download_known(), g['ongrid_known'])
    hist = e.refine(LABELS_BADER, 'all', 1, g['dist_mat'], g['T_grad'])
    assert hist == [(int(g['ongrid_edges']), int(g['ongrid_trace1_changed']))]
    np.testing.assert_array_equal(e.download_labels(LABELS_BADER, g['ongrid_trace1_volumes'].dtype),
                                  g['ongrid_trace1_volumes'])
    np.testing.assert_array_equal(e.download_known(), g['ongrid_trace1_known'])
    e.close()


@pytest.mark.parametrize('tag,mode', [('changed3', ('changed', 3)), ('all_inf', ('all', -1)),
                                      ('all2', ('all', 2))])
def test_refine_driver_golden_bit_exact(th, golden, tag, mode):
    g = golden
    v = g['ongrid_volumes'].copy()
    th.refine('neargrid', mode, g['rho'], v, g['dist_mat'], g['T_grad'], 1)
    np.testing.assert_array_equal(v, g[f'ongrid_refine_{tag}'])


def test_edge_check_known_golden_bit_exact(golden):
    """after trace iteration 1, the 'changed'-mode reclassification"""
    from pybader_b200.engine import Engine, LABELS_BADER
    from oracle import pyoracle as orc
    g = golden
    # reference state after two 'changed' iterations == oracle (pinned to golden)
    v = g['ongrid_volumes'].astype(np.int32)
    orc.refine('neargrid', ('changed', 2), g['rho'], v, g['dist_mat'], g['T_grad'])
    e = Engine(g['rho'].shape)
    e.upload_density(0, g['rho'])
    e.upload_labels(LABELS_BADER, g['ongrid_volumes'])
    hist = e.refine(LABELS_BADER, 'changed', 2, g['dist_mat'], g['T_grad'])
    np.testing.assert_array_equal(e.download_labels(LABELS_BADER, np.int32), v)
    assert hist[0] == (int(g['ongrid_edges']), int(g['ongrid_trace1_changed']))
    if len(hist) > 1:
        assert hist[1][0] == int(g['ongrid_check_counts'][1])
    e.close()


def test_refine_unknown_method_is_noop(th, golden):
    g = golden
    v = g['ongrid_volumes'].copy()
    th.refine('ongrid', ('all', -1), g['rho'], v, g['dist_mat'], g['T_grad'], 1)
    np.testing.assert_array_equal(v, g['ongrid_volumes'])
    th.refine('neargrid', ('all', 0), g['rho'], v, g['dist_mat'], g['T_grad'], 1)
    np.testing.assert_array_equal(v, g['ongrid_volumes'])
    with pytest.raises(AttributeError):
        th.bader_calc('weight', g['rho'], v, g['dist_mat'], g['T_grad'], 1)


def check_neargrid(labels, maxima, ref_labels, ref_maxima, rho, orc, what):
    a = canonical(labels, maxima)
    b = canonical(ref_labels, ref_maxima)
    diff = a != b
    n = labels.size
    agree = 1.0 - diff.sum() / n
    assert agree >= LABEL_AGREE, f"{what}: only {agree:.6f} of labels agree"
    if diff.any():
        known = np.zeros(labels.shape, dtype=np.int8)
        orc.edge_find(known, rho, ref_labels)
        assert np.all(known[diff] < 0), f"{what}: a disagreement lies off the Bader surfaces"
    return int(diff.sum())


def test_neargrid_golden(th, ut, orc, golden):
    g = golden
    mx, vol = th.bader_calc('neargrid', g['rho'], fresh_volumes(ut, g), g['dist_mat'], g['T_grad'], 1)
    th.refine('neargrid', ('changed', 2), g['rho'], vol, g['dist_mat'], g['T_grad'], 1)
    # bader_calc(neargrid) leaves the labels quiescent under incremental
    # propagation; the full pass of refine() may still move a handful of voxels
    assert th.refine.last_history[0][1] <= max(2, 1e-4 * vol.size)
    # same set of maxima
    key = lambda m: sorted(map(tuple, m.tolist()))
    assert key(mx) == key(g['neargrid_maxima'])
    check_neargrid(vol, mx, g['neargrid_refine_changed2'], g['neargrid_maxima'], g['rho'], orc,
                   g['name'])
    # charges per volume, matched through the maxima
    n = mx.shape[0]
    dV = float(g['voxel_volume'])
    q, v = np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, dV, g['rho'], vol)
    order = {tuple(m): i for i, m in enumerate(g['neargrid_maxima'].tolist())}
    perm = np.array([order[tuple(m)] for m in mx.tolist()])
    np.testing.assert_allclose(q, g['bader_charge'][perm], rtol=REL_TOL, atol=1e-9)
    np.testing.assert_allclose(v, g['bader_volume'][perm], rtol=REL_TOL)


def test_sums_atoms_surface_golden(th, ut, golden):
    """charge_sum / assign_to_atoms / surface_distance on the reference's own labels"""
    from pybader_b200 import geometry as geo
    g = golden
    final = g['neargrid_refine_changed2']
    n = g['neargrid_maxima'].shape[0]
    dV = float(g['voxel_volume'])
    q, v = np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, dV, g['rho'], final)
    np.testing.assert_allclose(q, g['bader_charge'], rtol=1e-12)
    np.testing.assert_allclose(v, g['bader_volume'], rtol=1e-9)
    if 'spin' in g:
        s, v2 = np.zeros(n), np.zeros(n)
        ut.charge_sum(s, v2, dV, g['spin'], final)
        np.testing.assert_allclose(s, g['bader_spin'], rtol=1e-10, atol=1e-13)
    ba, bd, av = th.assign_to_atoms(g['bader_maxima_cart'], g['atoms'], g['lattice'], final, 1)
    np.testing.assert_array_equal(ba, g['bader_atoms'])
    np.testing.assert_allclose(bd, g['bader_distance'], rtol=1e-14)
    assert av.dtype == g['atoms_volumes'].dtype
    np.testing.assert_array_equal(av, g['atoms_volumes'])
    na = g['atoms'].shape[0]
    q, v = np.zeros(na), np.zeros(na)
    ut.charge_sum(q, v, dV, g['rho'], av)
    np.testing.assert_allclose(q, g['atoms_charge'], rtol=1e-12)
    np.testing.assert_allclose(v, g['atoms_volume'], rtol=1e-9)
    off = np.dot(g['voxel_offset'], geo.voxel_lattice(g['lattice'], g['rho'].shape))
    sd = th.surface_distance(g['rho'], av, g['lattice'], g['atoms'] - off, 1)
    sd = np.zeros(na) if sd is None else sd
    np.testing.assert_allclose(sd, g['atoms_surface_distance'], rtol=1e-14)
    m = ut.volume_mask(av, g['rho'], 0)
    np.testing.assert_array_equal(m, np.where(av == 0, g['rho'], 0.0))


# ------------------------------------------------------- oracle, seeded ----
def seeded_cases():
    from pybader_b200 import synth
    c = synth.case_c1(48)
    yield 'c1_48', c, None
    c = synth.case_rocksalt(40, cells=2, offset=0.13, a=5.64)
    yield 'rocksalt40', c, None
    c = synth.case_triclinic((36, 40, 44), n_atoms=8, seed=3)
    c['sigmas'] = c['sigmas'] * 2.5
    yield 'triclinic_vac', c, 1e-3
    c = synth.case_slab((17, 33, 70), n_atoms=5, seed=5)      # ragged: no tile multiple
    c['sigmas'] = c['sigmas'] * 2.0
    yield 'ragged_slab', c, 2e-2


@pytest.fixture(scope='module', params=list(seeded_cases()), ids=lambda p: p[0])
def seeded(request):
    from pybader_b200 import geometry as geo, synth
    name, c, tol = request.param
    rho, atoms = synth.make(c)
    return dict(name=name, rho=rho, atoms=atoms, lattice=c['lattice'], vacuum_tol=tol,
                dist_mat=geo.distance_matrix(c['lattice'], rho.shape),
                T_grad=geo.T_grad(c['lattice'], rho.shape),
                voxel_volume=geo.voxel_volume(c['lattice'], rho.shape))


def oracle_fresh(orc, s):
    v = np.zeros(s['rho'].shape, dtype=np.int32)
    if s['vacuum_tol'] is not None:
        orc.vacuum_assign(s['rho'], v, s['vacuum_tol'], s['rho'], s['voxel_volume'])
    return v


def gpu_fresh(ut, s):
    v = np.zeros(s['rho'].shape, dtype=np.int32)
    if s['vacuum_tol'] is not None:
        ut.vacuum_assign(s['rho'], v, s['vacuum_tol'], s['rho'], s['voxel_volume'])
    return v


def test_ongrid_vs_oracle_bit_exact(th, ut, orc, seeded):
    s = seeded
    v0 = gpu_fresh(ut, s)
    np.testing.assert_array_equal(v0, oracle_fresh(orc, s))
    mx, vol = th.bader_calc('ongrid', s['rho'], v0, s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('ongrid', s['rho'], oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    np.testing.assert_array_equal(mx, rmx)
    assert vol.dtype == rvol.dtype
    np.testing.assert_array_equal(vol, rvol)


@pytest.mark.parametrize('mode', [('changed', 3), ('changed', -1), ('all', -1)])
def test_ongrid_refine_vs_oracle_bit_exact(th, orc, seeded, mode):
    s = seeded
    if mode == ('changed', -1) and s['vacuum_tol'] is not None:
        pytest.skip("'changed' to convergence floods the vacuum in the reference (SURVEY A.5)")
    _, seed = orc.bader_calc('ongrid', s['rho'], oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    a, b = seed.copy(), seed.copy()
    log = []
    orc.refine('neargrid', mode, s['rho'], a, s['dist_mat'], s['T_grad'], log=log)
    th.refine('neargrid', mode, s['rho'], b, s['dist_mat'], s['T_grad'], 1)
    np.testing.assert_array_equal(b, a)
    assert [c for _, c in th.refine.last_history][:len(log)] == [c for _, c in log]


@pytest.mark.parametrize('mode', [('changed', 2), ('all', 2)])
def test_neargrid_vs_oracle(th, ut, orc, seeded, mode):
    s = seeded
    mx, vol = th.bader_calc('neargrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', mode, s['rho'], vol, s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('neargrid', s['rho'], oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    orc.refine('neargrid', mode, s['rho'], rvol, s['dist_mat'], s['T_grad'])
    key = lambda m: sorted(map(tuple, m.tolist()))
    assert key(mx) == key(rmx)
    order = {tuple(m): i for i, m in enumerate(rmx.tolist())}
    perm = np.array([order[tuple(m)] for m in mx.tolist()])
    n = mx.shape[0]
    # The reference's 'changed' mode relabels a few VACUUM voxels next to edges
    # that moved in its first iteration (edge_check does not test for vacuum,
    # refinement.py:446-448).  Which ones depends on its
    # scan-order dependent raw labels, so they cannot be reproduced; they are
    # counted and credited explicitly instead of hiding them in the tolerance.
    quirk = (vol == -1) & (rvol >= 0)
    if mode[0] == 'all' or s['vacuum_tol'] is None:
        assert not quirk.any()
    assert quirk.sum() <= max(3, 1e-4 * vol.size)
    vol_cmp = vol.copy()
    inv = np.argsort(perm)
    vol_cmp[quirk] = inv[rvol[quirk]]
    ndiff = check_neargrid(vol_cmp, mx, rvol, rmx, s['rho'], orc, s['name'])
    q, v, rq, rv = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, s['voxel_volume'], s['rho'], vol_cmp)
    orc.charge_sum(rq, rv, s['voxel_volume'], s['rho'], rvol)
    np.testing.assert_allclose(q, rq[perm], rtol=REL_TOL, atol=1e-9)
    np.testing.assert_allclose(v, rv[perm], rtol=REL_TOL)
    # numbering: volume numbers ascend with each volume's first voxel
    first = [int(np.flatnonzero(vol.ravel() == k)[0]) for k in range(n)]
    assert first == sorted(first)
    print(f"{s['name']} {mode}: {ndiff} of {vol.size} voxels differ from the reference path, "
          f"{int(quirk.sum())} vacuum voxels relabelled by the reference only")


def test_fp32_seed_equals_exact_seed(th, ut, seeded, monkeypatch):
    """bader_calc('neargrid') seeds its rounds with the fp32-ranked stencil (csrc/seed.cuh);
    BDR_SEED_EXACT=1 seeds them with the bit-exact ongrid pointers instead.  Same maxima in
    the same order, and labels that agree like two runs of the reference do (>= 99.9 %,
    differences on edge voxels only)."""
    s = seeded
    mx, vol = th.bader_calc('neargrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', ('changed', 2), s['rho'], vol, s['dist_mat'], s['T_grad'], 1)
    monkeypatch.setenv('BDR_SEED_EXACT', '1')
    mx2, vol2 = th.bader_calc('neargrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', ('changed', 2), s['rho'], vol2, s['dist_mat'], s['T_grad'], 1)
    np.testing.assert_array_equal(mx, mx2)
    assert np.mean(vol == vol2) >= 0.999
    assert np.array_equal(vol == -1, vol2 == -1)


@pytest.mark.parametrize('shape', [(40, 36, 70), (33, 20, 129), (16, 48, 64)])
def test_edge_pass_equality_bits_match_minmax_kernel(shape, monkeypatch):
    """refinement.edge_find through label-equality bits (csrc/edge.cuh) against the min/max
    streaming kernel it replaced (BDR_EDGE_OLD=1): identical `known` and edge count, on
    ragged shapes (partial words, nz not a multiple of 4) with and without vacuum"""
    from pybader_b200.engine import Engine, LABELS_BADER
    rng = np.random.default_rng(sum(shape))
    # blocky random labels (edges everywhere), a vacuum slab and scattered vacuum voxels
    coarse = rng.integers(0, 5, size=tuple((n + 5) // 6 for n in shape))
    lab = np.kron(coarse, np.ones((6, 6, 6), dtype=np.int64))[:shape[0], :shape[1], :shape[2]]
    lab = np.ascontiguousarray(lab, dtype=np.int32)
    rho = rng.random(shape)
    for vac in (False, True):
        l = lab.copy()
        if vac:
            l[:, :, shape[2] // 2: shape[2] // 2 + 9] = -1
            l[rng.random(shape) < 0.02] = -1
            l[0, 0, 0] = -1
            l[-1, -1, -1] = -1
        out = []
        for old in (False, True):
            if old:
                monkeypatch.setenv('BDR_EDGE_OLD', '1')
            else:
                monkeypatch.delenv('BDR_EDGE_OLD', raising=False)
            e = Engine(shape)
            e.upload_density(0, rho)
            e.upload_labels(LABELS_BADER, l)
            n = e.edge_find(LABELS_BADER)
            out.append((n, e.download_known()))
            e.close()
        assert out[0][0] == out[1][0]
        np.testing.assert_array_equal(out[0][1], out[1][1])


# ------------------------------ bader-read's re-threshold flow ----
def test_rethreshold_labelled_volumes(th, ut, orc, seeded):
    """entry_points.py:238-255: a finished run is re-thresholded with a larger vacuum_tol,
    `volumes_init(volumes=bader_volumes)` + `sum_volumes`: vacuum_assign on an array that
    already holds labels, then charge_sum -- identical to the oracle"""
    s = seeded
    _, vol = th.bader_calc('ongrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    n = int(vol.max()) + 1
    tol = float(np.quantile(s['rho'], 0.35))
    a, b = vol.copy(), vol.copy()
    ra, rq, rv = orc.vacuum_assign(s['rho'], a, tol, s['rho'], s['voxel_volume'])
    gb, gq, gv = ut.vacuum_assign(s['rho'], b, tol, s['rho'], s['voxel_volume'])
    assert gb.dtype == vol.dtype
    np.testing.assert_array_equal(gb, ra)
    assert (gb == -1).sum() >= (vol == -1).sum() and (gb == -1).sum() >= 0.3 * gb.size
    assert gq == pytest.approx(rq, rel=1e-12) and gv == pytest.approx(rv, rel=1e-12)
    q, v, oq, ov = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, s['voxel_volume'], s['rho'], gb)
    orc.charge_sum(oq, ov, s['voxel_volume'], s['rho'], ra)
    np.testing.assert_allclose(q, oq, rtol=1e-12)
    np.testing.assert_allclose(v, ov, rtol=1e-12)


# ------------------------------------------------------- noisy densities ----
@pytest.mark.parametrize('noise', [1.0, 3.0])
def test_noisy_density_many_small_volumes(th, ut, orc, noise):
    """a density with multiplicative noise has hundreds of tiny maxima and few-voxel
    volumes (outside the BASELINE configs, which are smooth): ongrid stays bit-exact;
    neargrid finds the same maxima, and with both sides refined to convergence
    (('all', -1): on such data the reference is still moving after 2 iterations) the
    labels agree on >= 99.9 % of the voxels"""
    from pybader_b200 import geometry as geo, synth
    c = synth.case_c1(40)
    rho, _ = synth.make(c)
    rng = np.random.default_rng(int(noise * 100))
    rho = np.ascontiguousarray(rho * (1.0 + noise * rng.standard_normal(rho.shape)).clip(0.05))
    s = dict(name=f'noisy{noise}', rho=rho, vacuum_tol=None, lattice=c['lattice'],
             dist_mat=geo.distance_matrix(c['lattice'], rho.shape),
             T_grad=geo.T_grad(c['lattice'], rho.shape),
             voxel_volume=geo.voxel_volume(c['lattice'], rho.shape))
    mx, vol = th.bader_calc('ongrid', rho, gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('ongrid', rho, oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    np.testing.assert_array_equal(mx, rmx)
    np.testing.assert_array_equal(vol, rvol)
    assert mx.shape[0] > 50
    mx, vol = th.bader_calc('neargrid', rho, gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', ('all', -1), rho, vol, s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('neargrid', rho, oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    orc.refine('neargrid', ('all', -1), rho, rvol, s['dist_mat'], s['T_grad'])
    key = lambda m: sorted(map(tuple, m.tolist()))
    assert key(mx) == key(rmx)
    ndiff = int((canonical(vol, mx) != canonical(rvol, rmx)).sum())
    assert ndiff <= 1e-3 * vol.size      # measured on B200: 7 and 5 of 64,000
    print(f"{s['name']}: {mx.shape[0]} maxima, {ndiff} of {vol.size} voxels differ from the reference path")


# ---------------------------------------- BASELINE configs 1 and 2, full size ----
def _case_dict(name, c, tol=None):
    from pybader_b200 import geometry as geo, synth
    # orthorhombic cells factorise: 256^3 from 1-D tables in a fraction of a second
    tx, ty, tz = synth.separable_tables(c)
    rho = np.ascontiguousarray(np.einsum('ai,aj,ak->ijk', tx, ty, tz, optimize=True))
    atoms = c['frac_atoms'] @ c['lattice']
    return dict(name=name, rho=rho, atoms=atoms, lattice=c['lattice'], vacuum_tol=tol,
                dist_mat=geo.distance_matrix(c['lattice'], rho.shape),
                T_grad=geo.T_grad(c['lattice'], rho.shape),
                voxel_volume=geo.voxel_volume(c['lattice'], rho.shape))


def test_config1_full_size_vs_oracle(th, ut, orc):
    """BASELINE config 1: 3-atom cubic cell 96^3, neargrid + refine ('changed', 2), against the
    (reference-pinned) oracle: labels >= 99.9 % identical with differences on Bader surfaces
    only, same maxima, charges and volumes within 1e-6"""
    from pybader_b200 import synth
    s = _case_dict('c1_96', synth.case_c1(96))
    mx, vol = th.bader_calc('neargrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', ('changed', 2), s['rho'], vol, s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('neargrid', s['rho'], oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    orc.refine('neargrid', ('changed', 2), s['rho'], rvol, s['dist_mat'], s['T_grad'])
    key = lambda m: sorted(map(tuple, m.tolist()))
    assert key(mx) == key(rmx)
    ndiff = check_neargrid(vol, mx, rvol, rmx, s['rho'], orc, s['name'])
    order = {tuple(m): i for i, m in enumerate(rmx.tolist())}
    perm = np.array([order[tuple(m)] for m in mx.tolist()])
    n = mx.shape[0]
    q, v, rq, rv = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, s['voxel_volume'], s['rho'], vol)
    orc.charge_sum(rq, rv, s['voxel_volume'], s['rho'], rvol)
    np.testing.assert_allclose(q, rq[perm], rtol=REL_TOL)
    np.testing.assert_allclose(v, rv[perm], rtol=REL_TOL)
    print(f"config 1 (96^3): {ndiff} of {vol.size} voxels differ from the reference path")


def test_config2_full_size_ongrid_bit_exact(th, ut, orc):
    """BASELINE config 2: rocksalt-like 64-atom cell 256^3, method=ongrid: maxima list, labels
    and label dtype bit-identical to the oracle (tie-breaking included)"""
    from pybader_b200 import synth
    s = _case_dict('rocksalt256', synth.case_rocksalt(256, cells=4, offset=0.13))
    mx, vol = th.bader_calc('ongrid', s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    rmx, rvol = orc.bader_calc('ongrid', s['rho'], oracle_fresh(orc, s), s['dist_mat'], s['T_grad'])
    np.testing.assert_array_equal(mx, rmx)
    assert vol.dtype == rvol.dtype
    np.testing.assert_array_equal(vol, rvol)


def test_config3_full_size_vs_oracle(th, ut, orc):
    """BASELINE config 3 at FULL size (triclinic 128-atom cell 360x360x480, vacuum_tol 1e-3,
    neargrid + refine ('changed', 2)) against the reference-pinned oracle on the same bytes
    (the density is generated on the device and downloaded; ~1 min of oracle on one host
    core).  Same bars and the same vacuum-quirk accounting as test_neargrid_vs_oracle."""
    from pybader_b200 import geometry as geo, session, synth
    from pybader_b200.engine import Engine
    session.close_all()
    c = synth.case_triclinic((360, 360, 480), n_atoms=128, seed=1234)
    shape = c['shape']
    e = Engine(shape)
    e.synth_general(0, c['lattice'], c['frac_atoms'], c['amps'], c['sigmas'])
    rho = e.download_density(0)
    e.close()
    s = dict(name='c3_full', rho=rho, lattice=c['lattice'], vacuum_tol=1e-3,
             dist_mat=geo.distance_matrix(c['lattice'], shape), T_grad=geo.T_grad(c['lattice'], shape),
             voxel_volume=geo.voxel_volume(c['lattice'], shape))
    mode = ('changed', 2)
    v0 = gpu_fresh(ut, s)
    r0 = oracle_fresh(orc, s)
    np.testing.assert_array_equal(v0, r0)
    mx, vol = th.bader_calc('neargrid', rho, v0, s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', mode, rho, vol, s['dist_mat'], s['T_grad'], 1)
    hist = list(th.refine.last_history)
    rmx, rvol = orc.bader_calc('neargrid', rho, r0, s['dist_mat'], s['T_grad'])
    orc.refine('neargrid', mode, rho, rvol, s['dist_mat'], s['T_grad'])
    key = lambda m: sorted(map(tuple, m.tolist()))
    assert key(mx) == key(rmx)
    assert vol.dtype == rvol.dtype                      # int16: >= 128 maxima
    order = {tuple(m): i for i, m in enumerate(rmx.tolist())}
    perm = np.array([order[tuple(m)] for m in mx.tolist()])
    n = mx.shape[0]
    quirk = (vol == -1) & (rvol >= 0)                   # SURVEY A.5, see test_neargrid_vs_oracle
    assert quirk.sum() <= 1e-4 * vol.size              # measured: 851 of 62.2 M (1.4e-5)
    # ... and this engine's edge_check hands over the vacuum voxels next to ITS changed voxels
    mine_only = (vol >= 0) & (rvol == -1)
    assert mine_only.sum() <= 1e-4 * vol.size
    vol_cmp = vol.copy()
    vol_cmp[quirk] = np.argsort(perm)[rvol[quirk]]
    vol_cmp[mine_only] = -1
    ndiff = check_neargrid(vol_cmp, mx, rvol, rmx, rho, orc, s['name'])
    q, v, rq, rv = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, s['voxel_volume'], rho, vol_cmp)
    orc.charge_sum(rq, rv, s['voxel_volume'], rho, rvol)
    np.testing.assert_allclose(q, rq[perm], rtol=REL_TOL, atol=1e-9)
    np.testing.assert_allclose(v, rv[perm], rtol=REL_TOL)
    print(f"config 3 (360x360x480, {n} maxima): {ndiff} of {vol.size} voxels differ from the reference "
          f"path, {int(quirk.sum())} vacuum voxels relabelled by the reference only, "
          f"{int(mine_only.sum())} by this engine only; refine history {hist}")
    session.close_all()


def test_raw_neargrid_labels_vs_reference_raw_labels(th, ut):
    """bader_calc('neargrid') WITHOUT refinement against the real reference's
    raw (scan-order dependent) labels on BASELINE config 1.  The reference's raw labels differ
    from its own refined labels on 0.1-0.3 % of voxels (all on Bader surfaces); this
    engine's bader_calc returns the refinement fixed point, so its agreement with the raw labels
    is bounded by the same figure.  Measured and reported here; the 99.9 % bar of the north star
    is met against the refined result (test_config1_* in test_config1_reference.py)."""
    import hashlib
    import os
    from pybader_b200 import geometry as geo, synth
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_c1', 'c1_96.npz')
    with np.load(path) as z:
        g = {k: z[k] for k in z.files}
    c = synth.case_c1(96)
    rho, _ = synth.make(c)
    if not np.array_equal(np.frombuffer(hashlib.sha256(rho.tobytes()).digest(), dtype=np.uint8),
                          g['rho_sha256']):
        pytest.skip("this CPU's exp() gives other density bytes than the golden run")
    dist, T = geo.distance_matrix(c['lattice'], rho.shape), geo.T_grad(c['lattice'], rho.shape)
    mx, vol = th.bader_calc('neargrid', rho, np.zeros(rho.shape, np.int32), dist, T, 1)
    np.testing.assert_array_equal(mx, g['neargrid_maxima'])
    raw = g['neargrid_raw_labels']
    refined = g['neargrid_refined_labels']
    d_raw = int(np.count_nonzero(vol != raw))
    d_ref = int(np.count_nonzero(vol != refined))
    own = int(np.count_nonzero(raw != refined))
    print(f"raw bader_calc('neargrid') on config 1: {d_raw} of {vol.size} voxels differ from the "
          f"reference's raw labels ({1 - d_raw / vol.size:.6f} agree), {d_ref} from its refined labels; "
          f"the reference's raw and refined labels differ on {own}")
    assert d_ref <= 1e-3 * vol.size
    assert d_raw <= own + 1e-3 * vol.size      # no further from the raw labels than refinement itself


# ------------------------------------------------ properties at size -------
def test_properties_256(th, ut):
    """size-independent properties on a 256^3 rocksalt cell (BASELINE config 2
    shape), device-generated input"""
    from pybader_b200 import geometry as geo, synth
    from pybader_b200.engine import Engine, LABELS_BADER
    n = 256
    c = synth.case_rocksalt(n, cells=4, offset=0.13)
    e = Engine((n, n, n))
    tabs = synth.separable_tables(c)
    e.synth_separable(0, *tabs)
    dist = geo.distance_matrix(c['lattice'], (n, n, n))
    T = geo.T_grad(c['lattice'], (n, n, n))
    dV = geo.voxel_volume(c['lattice'], (n, n, n))
    e.clear_labels()
    mx = e.bader_calc('ongrid', dist, T)
    assert mx.shape[0] == 64
    lab = e.download_labels(LABELS_BADER, np.int32)
    assert lab.min() == 0 and lab.max() == 63
    # every maximum carries its own number; numbering ascends with first voxel
    assert [int(lab[tuple(m)]) for m in mx] == list(range(64))
    flat = lab.ravel()
    first = np.full(64, flat.size, dtype=np.int64)
    np.minimum.at(first, flat, np.arange(flat.size))
    assert np.all(np.diff(first) > 0)
    # charge conservation: sum over volumes == total
    rho = e.download_density(0)
    q, v = np.zeros(64), np.zeros(64)
    e.charge_sum(LABELS_BADER, 0, dV, q, v)
    assert q.sum() == pytest.approx(rho.sum() * dV, rel=1e-12)
    assert v.sum() == pytest.approx(flat.size * dV, rel=1e-12)
    # neargrid fixed point is idempotent under another refinement
    e.clear_labels()
    mxn = e.bader_calc('neargrid', dist, T)
    assert sorted(map(tuple, mxn.tolist())) == sorted(map(tuple, mx.tolist()))
    hist = e.refine(LABELS_BADER, 'all', -1, dist, T)
    assert hist[0][1] <= 1e-5 * flat.size
    assert hist[-1][1] == 0
    # with the fixed point certified inside bader_calc, refine finds nothing to do
    e.set_option(0, 1)
    e.clear_labels()
    e.bader_calc('neargrid', dist, T)
    lab_fp = e.download_labels(LABELS_BADER, np.int32)
    hist = e.refine(LABELS_BADER, 'all', -1, dist, T)
    assert hist[0][1] == 0
    np.testing.assert_array_equal(e.download_labels(LABELS_BADER, np.int32), lab_fp)
    e.close()


def test_shared_reciprocal_division_is_ieee():
    """the trace kernel's shared-reciprocal division == hardware fp64 division"""
    import ctypes
    from pybader_b200.engine import Engine
    from pybader_b200._lib import check
    e = Engine((4, 4, 4))
    bad = ctypes.c_int64(-1)
    for seed in (1, 2, 3):
        check(e.lib.bdr_selftest_div(e.h, 1 << 24, seed, ctypes.byref(bad)))
        assert bad.value == 0
    e.close()


@pytest.mark.parametrize('method,mode', [('ongrid', ('all', 0)), ('neargrid', ('changed', 2))])
def test_one_shot_run_equals_staged(th, ut, seeded, method, mode):
    """bdr_run (host buffers in, pipelined upload + stencil, host buffers out)
    gives exactly what the stage-by-stage entry points give"""
    from pybader_b200.engine import Engine
    s = seeded
    mx, vol = th.bader_calc(method, s['rho'], gpu_fresh(ut, s), s['dist_mat'], s['T_grad'], 1)
    th.refine('neargrid', mode, s['rho'], vol, s['dist_mat'], s['T_grad'], 1)
    n = mx.shape[0]
    q, v = np.zeros(n), np.zeros(n)
    ut.charge_sum(q, v, s['voxel_volume'], s['rho'], vol)
    e = Engine(s['rho'].shape)
    lab, mx2, q2, v2 = e.run(s['rho'], s['vacuum_tol'], s['voxel_volume'], method, mode[0], mode[1],
                             s['dist_mat'], s['T_grad'], label_dtype=vol.dtype)
    e.close()
    np.testing.assert_array_equal(mx2, mx)
    np.testing.assert_array_equal(lab, vol)
    np.testing.assert_allclose(q2, q, rtol=1e-12)
    np.testing.assert_allclose(v2, v, rtol=1e-12)


# ------------------------------------- BASELINE configs 3 and 4 at full size ----
def _properties(e, lab, mx, dV, vac_count=0, check_order=True):
    from pybader_b200.engine import LABELS_BADER
    n = mx.shape[0]
    flat = lab.ravel()
    assert flat.min() == (-1 if vac_count else 0) and flat.max() == n - 1
    assert int((flat == -1).sum()) == vac_count
    # every maximum carries its own number; numbers ascend with the first voxel
    assert [int(lab[tuple(m)]) for m in mx] == list(range(n))
    if check_order:
        first = np.full(n, flat.size, dtype=np.int64)
        step = 1 << 24
        for lo in range(0, flat.size, step):
            vals, idx = np.unique(flat[lo:lo + step], return_index=True)
            keep = vals >= 0
            np.minimum.at(first, vals[keep], idx[keep] + lo)
        assert np.all(np.diff(first) > 0)
    q, v = np.zeros(n), np.zeros(n)
    e.charge_sum(LABELS_BADER, 0, dV, q, v)
    return q, v


def test_config3_triclinic_vacuum_full_size():
    """BASELINE config 3: triclinic 128-atom cell 360x360x480, vacuum_tol 1e-3,
    neargrid + refine ('changed', 2): size-independent properties"""
    from pybader_b200 import geometry as geo, synth
    from pybader_b200.engine import Engine, LABELS_BADER
    c = synth.case_triclinic((360, 360, 480), n_atoms=128, seed=1234)
    shape = c['shape']
    e = Engine(shape)
    e.synth_general(0, c['lattice'], c['frac_atoms'], c['amps'], c['sigmas'])
    dist = geo.distance_matrix(c['lattice'], shape)
    T = geo.T_grad(c['lattice'], shape)
    dV = geo.voxel_volume(c['lattice'], shape)
    rho = e.download_density(0)
    tol = 1e-3
    nvac = int((rho <= tol).sum())
    assert 0 < nvac < rho.size
    e.clear_labels()
    vq, vv = e.vacuum_assign(tol, dV)
    assert vv == pytest.approx(nvac * dV, rel=1e-12)
    assert vq == pytest.approx(rho[rho <= tol].sum() * dV, rel=1e-9)
    mx_on = e.bader_calc('ongrid', dist, T)
    lab_on = e.download_labels(LABELS_BADER, np.int32)
    q_on, v_on = _properties(e, lab_on, mx_on, dV, nvac)
    assert q_on.sum() + vq == pytest.approx(rho.sum() * dV, rel=1e-10)
    # ongrid pointers are a pure function of the density: spot-check 400 voxels on the CPU
    rng = np.random.default_rng(7)
    W = np.array([[[dist[i, j, k] for k in (-1, 0, 1)] for j in (-1, 0, 1)] for i in (-1, 0, 1)])
    for x, y, z in rng.integers(0, shape, size=(400, 3)):
        if lab_on[x, y, z] < 0:
            continue
        p = (x, y, z)
        for _ in range(4000):
            rc, best, nxt = rho[p], rho[p], p
            for ix in (-1, 0, 1):
                for iy in (-1, 0, 1):
                    for iz in (-1, 0, 1):
                        q = ((p[0] + ix) % shape[0], (p[1] + iy) % shape[1], (p[2] + iz) % shape[2])
                        val = (rho[q] - rc) * W[ix + 1, iy + 1, iz + 1] + rc
                        if val > best:
                            best, nxt = val, q
            if nxt == p:
                break
            p = nxt
        assert lab_on[p] == lab_on[x, y, z] and tuple(mx_on[lab_on[p]]) == p
    # neargrid + refine: same maxima, quiescent, vacuum untouched by 'all' mode
    e.clear_labels()
    e.vacuum_assign(tol, dV)
    mx = e.bader_calc('neargrid', dist, T)
    assert sorted(map(tuple, mx.tolist())) == sorted(map(tuple, mx_on.tolist()))
    hist = e.refine(LABELS_BADER, 'changed', 2, dist, T)
    assert hist[0][1] <= 1e-5 * rho.size
    lab = e.download_labels(LABELS_BADER, np.int32)
    q, v = np.zeros(len(mx)), np.zeros(len(mx))
    e.charge_sum(LABELS_BADER, 0, dV, q, v)
    # 'changed' mode may hand a few vacuum voxels to volumes (reference quirk)
    moved = nvac - int((lab == -1).sum())
    assert 0 <= moved <= 1e-5 * rho.size
    assert v.sum() == pytest.approx((rho.size - nvac + moved) * dV, rel=1e-12)
    assert np.mean((lab >= 0) == (lab_on >= 0)) > 0.99999
    # neargrid and ongrid partitions agree away from the surfaces
    order = {tuple(m): i for i, m in enumerate(mx_on.tolist())}
    perm = np.array([order[tuple(m)] for m in mx.tolist()])
    same = np.where(lab >= 0, perm[np.maximum(lab, 0)], -1) == lab_on
    assert same.mean() > 0.97
    e.close()


def test_config4_slab_spin_full_size():
    """BASELINE config 4: 512x512x1024 slab with vacuum and a spin density"""
    from pybader_b200 import geometry as geo, synth
    from pybader_b200.engine import Engine, LABELS_ATOMS, LABELS_BADER
    c = synth.case_slab((512, 512, 1024), n_atoms=64, seed=4321)
    shape = c['shape']
    e = Engine(shape)
    tx, ty, tz = synth.separable_tables(c)
    e.synth_separable(0, tx, ty, tz)
    e.synth_separable(2, tx * c['spin_weights'][:, None], ty, tz)
    dist = geo.distance_matrix(c['lattice'], shape)
    T = geo.T_grad(c['lattice'], shape)
    dV = geo.voxel_volume(c['lattice'], shape)
    tol = 1e-3
    e.clear_labels()
    vq, vv = e.vacuum_assign(tol, dV)
    nvac = round(vv / dV)
    assert 0 < nvac < np.prod(shape)
    mx = e.bader_calc('neargrid', dist, T)
    hist = e.refine(LABELS_BADER, 'all', 2, dist, T)
    assert hist[0][1] <= 1e-5 * np.prod(shape) and hist[-1][1] <= hist[0][1]
    lab = e.download_labels(LABELS_BADER, np.int32)
    n = len(mx)
    assert 32 <= n <= 64          # close Gaussians merge into one maximum
    q, v = _properties(e, lab, mx, dV, nvac, check_order=False)
    assert v.sum() == pytest.approx((np.prod(shape) - nvac) * dV, rel=1e-12)
    s, v2 = np.zeros(n), np.zeros(n)
    e.charge_sum(LABELS_BADER, 2, dV, s, v2)
    np.testing.assert_array_equal(v2, v)
    # atoms: every maximum goes to its nearest atom; per-atom sums == sums of their volumes
    atoms = c['frac_atoms'] @ c['lattice']
    off = np.array([.5, .5, .5])                              # cube files, io/cube.py:154
    mcart = geo.maxima_fractional(mx, shape, off) @ c['lattice']
    who, d = e.assign_atoms(mcart, atoms, c['lattice'])
    assert who.min() >= 0 and who.max() < 64 and np.all(d >= 0)
    qa, va = np.zeros(64), np.zeros(64)
    e.charge_sum(LABELS_ATOMS, 0, dV, qa, va)
    np.testing.assert_allclose(qa, np.bincount(who, weights=q, minlength=64), rtol=1e-9)
    np.testing.assert_allclose(va, np.bincount(who, weights=v, minlength=64), rtol=1e-9)
    sa, _ = np.zeros(64), np.zeros(64)
    e.charge_sum(LABELS_ATOMS, 2, dV, sa, _)
    np.testing.assert_allclose(sa, np.bincount(who, weights=s, minlength=64), rtol=1e-7, atol=1e-9)
    sd = e.surface_distance(LABELS_ATOMS, c['lattice'], atoms)
    assert sd is not None and np.all(sd[np.bincount(who, minlength=64) > 0] > 0)
    e.close()
