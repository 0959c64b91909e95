"""numpy restatements of the two reformulations the CUDA path relies on, checked on the
CPU against the plain definitions (the kernels themselves are checked on the GPU):

* csrc/edge.cuh -- the label half of refinement.edge_find through equality bits:
  "27 labels all equal" as an AND of 26 adjacent-pair equalities, voxels with vacuum in
  reach classified by the definition;
* csrc/seed.cuh -- the fp32 "clearly uphill" acceptance rule implies the reference's exact
  fp64 ongrid criterion (rho_n - rho_c) * w + rho_c > rho_c for the accepted neighbour.
"""
import itertools

import numpy as np
import pytest

OFFS = [d for d in itertools.product((-1, 0, 1), repeat=3)]


def edge_candidates_definition(lab):
    """refinement.py:339-376 (label half): non-vacuum voxel with a non-vacuum neighbour of
    another label; vacuum neighbours are ignored"""
    out = np.zeros(lab.shape, dtype=bool)
    for d in OFFS:
        nb = np.roll(lab, tuple(-x for x in d), axis=(0, 1, 2))
        out |= (nb != -1) & (nb != lab)
    return out & (lab != -1)


def edge_candidates_eqbits(lab):
    """the formulation of csrc/edge.cuh"""
    eqz = lab == np.roll(lab, -1, axis=2)
    eqy = lab == np.roll(lab, -1, axis=1)
    eqx = lab == np.roll(lab, -1, axis=0)
    vac = lab == -1
    # per plane: the 3 x 3 (y, z) patch around (y, z) is uniform
    zrow = eqz & np.roll(eqz, 1, axis=2)                      # L[z-1] == L[z] == L[z+1]
    P = zrow & np.roll(zrow, 1, axis=1) & np.roll(zrow, -1, axis=1)
    P &= eqy & np.roll(eqy, 1, axis=1)                        # rows y-1, y, y+1 joined at column z
    uniform = P & np.roll(P, 1, axis=0) & np.roll(P, -1, axis=0) & eqx & np.roll(eqx, 1, axis=0)
    cand = ~uniform & ~vac
    # vacuum within Chebyshev distance 1: the chain argument does not apply there
    vnear = np.zeros(lab.shape, dtype=bool)
    for d in OFFS:
        vnear |= np.roll(vac, d, axis=(0, 1, 2))
    deferred = cand & vnear
    cand &= ~deferred
    cand[deferred] = edge_candidates_definition(lab)[deferred]   # k_edge_deferred: by definition
    return cand, int(deferred.sum())


@pytest.mark.parametrize('shape,vacuum', [((9, 7, 40), False), ((12, 5, 33), True), ((3, 3, 3), True),
                                          ((2, 6, 70), False), ((16, 16, 16), True)])
def test_equality_bits_equal_definition(shape, vacuum):
    rng = np.random.default_rng(sum(shape))
    coarse = rng.integers(0, 4, size=tuple((n + 3) // 4 for n in shape))
    lab = np.kron(coarse, np.ones((4, 4, 4), dtype=np.int64))[:shape[0], :shape[1], :shape[2]].copy()
    if vacuum:
        lab[rng.random(shape) < 0.08] = -1
        lab[:, :, shape[2] // 2] = -1
    got, n_def = edge_candidates_eqbits(lab)
    np.testing.assert_array_equal(got, edge_candidates_definition(lab))
    assert (n_def > 0) == vacuum


def test_fp32_clearly_uphill_implies_exact_uphill():
    """seed.cuh accepts pair winner k when (max(f32(rho_a), f32(rho_b)) - f32(rho_c)) * w_k is at
    least |f32(rho_c)| * 2^-21 * w_max (and above an absolute floor); for that neighbour the
    exact fp64 expression of methods.py:110-112 must exceed rho_c"""
    rng = np.random.default_rng(11)
    w = rng.uniform(0.5, 60.0, 13)
    wmax = w.max()
    c1 = np.float32(2.0 ** -21 * wmax * 1.0001)
    floor = np.float32(2.0 ** -96 * wmax)
    n = 400000
    for scale in (1.0, 1e-8, 1e-30, 1e12):
        rc = rng.lognormal(0, 2, n) * scale
        # neighbours from far below to a hair above the centre, incl. fp32-invisible differences
        rel = rng.choice([1e-16, 1e-12, 1e-9, 3e-8, 1e-7, 3e-7, 1e-6, 1e-3, 0.3], n) * rng.choice([-1, 1], n)
        rn = rc * (1.0 + rel * rng.random(n))
        k = rng.integers(0, 13, n)
        with np.errstate(over='ignore', under='ignore'):
            rcf, rnf = rc.astype(np.float32), rn.astype(np.float32)
            score = (rnf - rcf) * w[k].astype(np.float32)
            thr = np.maximum(np.abs(rcf) * c1, floor)
        accept = score >= thr
        exact = (rn - rc) * w[k] + rc            # numpy does not contract to FMA
        assert np.all(exact[accept] > rc[accept])
        assert np.all(rn[accept] > rc[accept])
        if scale in (1.0, 1e12):
            assert accept.sum() > 0.1 * n and (~accept).sum() > 0.1 * n


def seed_field_model(rho, dist_mat):
    """numpy model of k_seed_pointers (csrc/seed.cuh): pairs k / 26-k scored in fp32 with the
    index in the low mantissa bits, the winner accepted when clearly uphill, else the exact
    fp64 argmax of methods.py:87-117 (tests/shard_model.ongrid_pointers)"""
    from tests.shard_model import ongrid_pointers
    offs = [d for d in itertools.product((-1, 0, 1), repeat=3)]
    w = np.array([dist_mat[d] for d in offs])                   # negative indices: interface.py:249-258
    wf = w[:13].astype(np.float32)
    wmax = w[:13].max()
    c1, floor = np.float32(2.0 ** -21 * wmax * 1.0001), np.float32(2.0 ** -96 * wmax)
    rf = rho.astype(np.float32)
    lin = np.arange(rho.size, dtype=np.int64).reshape(rho.shape)
    best = np.zeros(rho.shape, dtype=np.float32)
    for k in range(13):
        a = np.roll(rf, tuple(-x for x in offs[k]), axis=(0, 1, 2))
        b = np.roll(rf, tuple(-x for x in offs[26 - k]), axis=(0, 1, 2))
        with np.errstate(over='ignore', under='ignore', invalid='ignore'):
            s = (np.maximum(a, b) - rf) * wf[k]
        tagged = ((s.view(np.uint32) & np.uint32(0xfffffff0)) | np.uint32(k)).view(np.float32)
        best = np.fmax(best, tagged)
    thr = np.maximum(np.abs(rf) * c1, floor)
    accept = best >= thr
    kk = (best.view(np.uint32) & np.uint32(15)).astype(np.int64)
    target = ongrid_pointers(rho, dist_mat)                      # exact fallback everywhere ...
    for k in range(13):                                          # ... overridden where accepted
        sel = accept & (kk == k)
        a = np.roll(rf, tuple(-x for x in offs[k]), axis=(0, 1, 2))
        b = np.roll(rf, tuple(-x for x in offs[26 - k]), axis=(0, 1, 2))
        la = np.roll(lin, tuple(-x for x in offs[k]), axis=(0, 1, 2))
        lb = np.roll(lin, tuple(-x for x in offs[26 - k]), axis=(0, 1, 2))
        target[sel] = np.where(a >= b, la, lb)[sel]
    return target, accept


@pytest.mark.parametrize('kind', ['smooth', 'noisy', 'quantised', 'tiny'])
def test_seed_field_is_ascending_with_the_exact_maxima(kind):
    """every seed pointer goes strictly uphill in the exact density (so the field is acyclic)
    and the fixed points are exactly the ongrid maxima"""
    from pybader_b200 import geometry as geo
    from tests.shard_model import ongrid_pointers
    rng = np.random.default_rng(len(kind))
    shape = (14, 11, 17)
    lattice = np.diag([3.0, 2.5, 4.0]) + rng.uniform(-0.3, 0.3, (3, 3))
    f = np.stack(np.meshgrid(*[np.arange(n) / n for n in shape], indexing='ij'), -1)
    rho = np.full(shape, 1e-3)
    for _ in range(4):
        c0, s, a = rng.random(3), rng.uniform(0.1, 0.3), rng.uniform(0.5, 2.0)
        d = (f - c0 + 0.5) % 1.0 - 0.5
        rho += a * np.exp(-(d ** 2).sum(-1) / (2 * s * s))
    if kind == 'noisy':
        rho *= 1.0 + 0.5 * rng.random(shape)
    if kind == 'quantised':
        rho = np.round(rho, 1) + 0.05
    if kind == 'tiny':
        rho *= 1e-44                                  # below fp32's normal range: all exact
    dist = geo.distance_matrix(lattice, shape)
    target, accept = seed_field_model(rho, dist)
    exact = ongrid_pointers(rho, dist)
    flat, lin = rho.reshape(-1), np.arange(rho.size)
    moved = target.reshape(-1) != lin
    assert np.all(flat[target.reshape(-1)[moved]] > flat[moved])
    np.testing.assert_array_equal(~moved, exact.reshape(-1) == lin)      # same maxima
    # where accepted, the chosen neighbour satisfies the reference's own criterion
    sel = accept.reshape(-1)
    assert np.all(sel <= moved)
    if kind in ('smooth', 'noisy'):
        assert sel.mean() > 0.9
    if kind == 'tiny':
        assert not sel.any()


# ---- round 2 reformulations ---------------------------------------------------------
def _eq_bits(lab):
    return (lab == np.roll(lab, -1, axis=2), lab == np.roll(lab, -1, axis=1),
            lab == np.roll(lab, -1, axis=0), lab == -1)


@pytest.mark.parametrize('shape', [(5, 6, 7), (3, 3, 3), (9, 4, 33)])
def test_equality_bits_patched_equal_recomputed(shape):
    """csrc/edge.cuh k_eq_update: after some voxels are relabelled, rewriting the seven bits
    around each of them (eqz at z-1 and z, eqy at y-1 and y, eqx at x-1 and x, the vacuum
    bit) gives exactly the bit volumes a full pass over the new labels computes; a
    renumbering (bijection on the labels >= 0) changes no bit at all"""
    rng = np.random.default_rng(sum(shape))
    lab = rng.integers(-1, 4, size=shape).astype(np.int32)
    eqz, eqy, eqx, vac = (b.copy() for b in _eq_bits(lab))
    new = lab.copy()
    changed = [tuple(int(v) for v in rng.integers(0, shape)) for _ in range(max(3, lab.size // 6))]
    for c in changed:
        new[c] = rng.integers(-1, 4)
    nx, ny, nz = shape
    for (x, y, z) in changed:                       # what one thread of k_eq_update does
        xm, xp, ym, yp, zm, zp = (x - 1) % nx, (x + 1) % nx, (y - 1) % ny, (y + 1) % ny, (z - 1) % nz, (z + 1) % nz
        l = new[x, y, z]
        eqz[x, y, z], eqz[x, y, zm] = l == new[x, y, zp], l == new[x, y, zm]
        eqy[x, y, z], eqy[x, ym, z] = l == new[x, yp, z], l == new[x, ym, z]
        eqx[x, y, z], eqx[xm, y, z] = l == new[xp, y, z], l == new[xm, y, z]
        vac[x, y, z] = l == -1
    for mine, full in zip((eqz, eqy, eqx, vac), _eq_bits(new)):
        np.testing.assert_array_equal(mine, full)
    perm = rng.permutation(4)
    renumbered = np.where(new >= 0, perm[np.maximum(new, 0)], -1)
    for a, b in zip(_eq_bits(new), _eq_bits(renumbered)):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('kind', ['smooth', 'noisy', 'quantised'])
@pytest.mark.parametrize('with_vacuum', [False, True])
def test_edge_maxima_are_ongrid_maxima(kind, with_vacuum):
    """premise of k_edge_confirm_roots: a voxel that refinement.edge_find
    keeps out of the edge set because it is a maximum (no non-vacuum neighbour of larger
    density, refinement.py:374-383) is one of the maxima methods.ongrid finds -- when the
    vacuum is a threshold of the same density.  So the density half of the exact edge pass
    only has to test the stencil pass's maxima.  Checked with the reference-pinned oracle."""
    from oracle import pyoracle as orc
    from pybader_b200 import geometry as geo
    rng = np.random.default_rng(11 + len(kind) + with_vacuum)
    shape = (10, 9, 12)
    lattice = np.diag([4.0, 3.5, 5.0]) + rng.uniform(-0.4, 0.4, (3, 3))
    f = np.stack(np.meshgrid(*[np.arange(n) / n for n in shape], indexing='ij'), -1)
    rho = np.full(shape, 1e-3)
    for _ in range(4):
        c0, s, a = rng.random(3), rng.uniform(0.12, 0.3), rng.uniform(0.5, 2.0)
        d = (f - c0 + 0.5) % 1.0 - 0.5
        rho += a * np.exp(-(d ** 2).sum(-1) / (2 * s * s))
    if kind == 'noisy':
        rho *= 1.0 + 0.3 * rng.random(shape)
    if kind == 'quantised':
        rho = np.round(rho, 1) + 0.05
    rho = np.ascontiguousarray(rho)
    dist, T = geo.distance_matrix(lattice, shape), geo.T_grad(lattice, shape)
    lab0 = np.zeros(shape, np.int32)
    if with_vacuum:
        lab0[rho <= np.quantile(rho, 0.3)] = -1
    mx, vol = orc.bader_calc('ongrid', rho, lab0.copy(), dist, T)
    roots = {tuple(m) for m in mx.tolist()}
    lab = vol.astype(np.int64)
    is_max = np.ones(shape, dtype=bool)
    for d in OFFS:
        nb_rho = np.roll(rho, tuple(-x for x in d), axis=(0, 1, 2))
        nb_lab = np.roll(lab, tuple(-x for x in d), axis=(0, 1, 2))
        is_max &= ~((nb_rho > rho) & (nb_lab != -1))
    edge_and_max = edge_candidates_definition(lab) & is_max
    _seen_edge_maxima.append(int(edge_and_max.sum()))
    for p in zip(*np.nonzero(edge_and_max)):
        assert tuple(int(v) for v in p) in roots, p


_seen_edge_maxima = []


def test_edge_maxima_cases_were_not_vacuous():
    assert sum(_seen_edge_maxima) > 0, "no case had a maximum on a Bader surface"
