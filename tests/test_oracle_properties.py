"""Property tests (hypothesis) of the oracle on small random grids and random triclinic
lattices, against independent numpy restatements and against invariants the reference
algorithm guarantees (SURVEY.md section 4, item ii).  CPU only."""
import itertools

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pyoracle as orc
from pybader_b200 import geometry as geo
from tests.shard_model import ongrid_pointers
from tests.test_algorithms_cpu import edge_candidates_definition


@st.composite
def small_case(draw):
    shape = (draw(st.integers(3, 9)), draw(st.integers(3, 9)), draw(st.integers(3, 12)))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    # a random triclinic cell that is far from degenerate
    lattice = np.diag(rng.uniform(2.0, 6.0, 3)) + rng.uniform(-0.6, 0.6, (3, 3))
    kind = draw(st.sampled_from(['smooth', 'noisy', 'quantised']))
    f = np.stack(np.meshgrid(*[np.arange(n) / n for n in shape], indexing='ij'), -1)
    rho = np.zeros(shape)
    for _ in range(draw(st.integers(1, 4))):
        c0, s, a = rng.random(3), rng.uniform(0.12, 0.4), rng.uniform(0.5, 2.0)
        d = (f - c0 + 0.5) % 1.0 - 0.5
        rho += a * np.exp(-(d ** 2).sum(-1) / (2 * s * s))
    rho += 1e-3
    if kind == 'noisy':
        rho *= 1.0 + 0.2 * rng.random(shape)
    if kind == 'quantised':
        rho = np.round(rho, 1) + 0.05            # plateaus and exact ties everywhere
    return np.ascontiguousarray(rho), lattice


@settings(max_examples=40, deadline=None, derandomize=True, database=None)
@given(small_case())
def test_ongrid_is_pointer_chasing_with_first_voxel_numbering(case):
    """methods.ongrid == chase the (ix,iy,iz)-ordered strict-'>' argmax pointer of every voxel
    to its fixed point; volumes numbered by their first voxel in C order; the k-th maximum
    carries label k"""
    rho, lattice = case
    dist = geo.distance_matrix(lattice, rho.shape)
    T = geo.T_grad(lattice, rho.shape)
    mx, vol = orc.bader_calc('ongrid', rho, np.zeros(rho.shape, np.int32), dist, T)
    ptr = ongrid_pointers(rho, dist).reshape(-1)
    root = ptr.copy()
    for _ in range(rho.size):
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    roots, first = np.unique(root, return_index=True)
    order = np.argsort(first)
    number = np.empty(rho.size, dtype=np.int64)
    number[roots[order]] = np.arange(len(roots))
    np.testing.assert_array_equal(vol.reshape(-1), number[root])
    np.testing.assert_array_equal(
        (mx[:, 0] * rho.shape[1] + mx[:, 1]) * rho.shape[2] + mx[:, 2], roots[order])


@settings(max_examples=40, deadline=None, derandomize=True, database=None)
@given(small_case(), st.booleans())
def test_edge_find_is_the_order_free_classification(case, with_vacuum):
    """refinement.edge_find == {edge and not a maximum: -2; Chebyshev-1 dilation of those: -1;
    other non-vacuum: 2; other vacuum: 0}, vacuum neighbours ignored"""
    rho, lattice = case
    dist = geo.distance_matrix(lattice, rho.shape)
    T = geo.T_grad(lattice, rho.shape)
    lab0 = np.zeros(rho.shape, np.int32)
    if with_vacuum:
        lab0[rho <= np.quantile(rho, 0.3)] = -1
    _, vol = orc.bader_calc('ongrid', rho, lab0, dist, T)
    lab = vol.astype(np.int64)
    known = np.zeros(rho.shape, dtype=np.int8)
    n_edges = orc.edge_find(known, rho, vol)
    cand = edge_candidates_definition(lab)
    is_max = np.ones(rho.shape, dtype=bool)
    for d in itertools.product((-1, 0, 1), repeat=3):
        nb_rho = np.roll(rho, tuple(-x for x in d), axis=(0, 1, 2))
        nb_lab = np.roll(lab, tuple(-x for x in d), axis=(0, 1, 2))
        is_max &= ~((nb_rho > rho) & (nb_lab != -1))
    edge = cand & ~is_max
    near = np.zeros(rho.shape, dtype=bool)
    for d in itertools.product((-1, 0, 1), repeat=3):
        near |= np.roll(edge, d, axis=(0, 1, 2))
    want = np.where(edge, -2, np.where(near, -1, np.where(lab == -1, 0, 2))).astype(np.int8)
    np.testing.assert_array_equal(known, want)
    assert n_edges == int(edge.sum())


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(small_case())
def test_refinement_invariants(case):
    """neargrid + refine to convergence: every voxel keeps a valid label, the k-th maximum is
    labelled k, charge is conserved, and one more 'all' pass changes nothing"""
    rho, lattice = case
    dist = geo.distance_matrix(lattice, rho.shape)
    T = geo.T_grad(lattice, rho.shape)
    dV = geo.voxel_volume(lattice, rho.shape)
    mx, vol = orc.bader_calc('neargrid', rho, np.zeros(rho.shape, np.int32), dist, T)
    orc.refine('neargrid', ('all', -1), rho, vol, dist, T)
    n = mx.shape[0]
    assert vol.min() >= 0 and vol.max() == n - 1
    assert [int(vol[tuple(m)]) for m in mx] == list(range(n))
    q, v = np.zeros(n), np.zeros(n)
    orc.charge_sum(q, v, dV, rho, vol)
    assert abs(q.sum() - rho.sum() * dV) <= 1e-10 * abs(rho.sum() * dV)
    assert abs(v.sum() - rho.size * dV) <= 1e-10 * rho.size * dV
    again = vol.copy()
    log = []
    orc.refine('neargrid', ('all', 1), rho, again, dist, T, log=log)
    np.testing.assert_array_equal(again, vol)
