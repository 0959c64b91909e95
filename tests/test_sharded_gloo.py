"""The sharded protocol (slab split, exit resolution, global numbering, halo
exchange, Jacobi refinement passes) on CPU over gloo with world_size 2 and 3,
against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(vac):
    from pybader_b200 import geometry as geo, synth
    if vac:
        c = synth.case_triclinic((48, 20, 22), n_atoms=6, seed=9)
        c['sigmas'] = c['sigmas'] * 2.5
        tol = 1e-3
    else:
        c = synth.case_rocksalt(48, cells=2, offset=0.13, a=5.64)
        c['shape'] = (48, 24, 20)
        tol = None
    rho, atoms = synth.make(c)
    shape = rho.shape
    return rho, tol, geo.distance_matrix(c['lattice'], shape), geo.T_grad(c['lattice'], shape), \
        geo.voxel_volume(c['lattice'], shape)


def _worker(rank, world, port, vac, halo, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pybader_b200.sharded import Comm, ShardedBader
        from tests.shard_model import ModelBackend
        rho, tol, dm, T, dV = _case(vac)
        comm = Comm()
        holder = {}

        def factory(window_shape, h):
            sb = holder['sb']
            return ModelBackend(window_shape, h, rho[sb.window_x], tol)

        sb = ShardedBader.__new__(ShardedBader)
        holder['sb'] = sb
        ShardedBader.__init__(sb, rho.shape, comm, factory, halo=halo)
        sb.backend.x0w, sb.backend.NX = sb.x0 - halo, rho.shape[0]
        mx = sb.ongrid(dm)
        lab_on = sb.owned_labels().numpy().copy()
        hist = sb.refine(dm, T, -1)
        lab_ng = sb.owned_labels().numpy().copy()
        n = mx.shape[0]
        q, v = sb.charge_sum(n, dV)
        # 'changed'-mode refinement across the slabs, from the same ongrid labels
        sb.backend.lab[...] = 0
        if tol is not None:
            sb.backend.lab[sb.backend.rho <= tol] = -1
        sb.ongrid(dm)
        hist_c = sb.refine(dm, T, 3, mode='changed')
        lab_c = sb.owned_labels().numpy().copy()
        np.savez(os.path.join(out, f'r{rank}.npz'), x0=sb.x0, x1=sb.x1, maxima=mx, lab_on=lab_on,
                 lab_ng=lab_ng, hist=np.array(hist), q=q, v=v, lab_c=lab_c, hist_c=np.array(hist_c))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,vac', [(2, False), (2, True), (3, False)])
def test_sharded_matches_single(tmp_path, world, vac):
    from oracle import pyoracle as orc
    halo = 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, vac, halo, str(tmp_path)), nprocs=world, join=True)
    rho, tol, dm, T, dV = _case(vac)
    v0 = np.zeros(rho.shape, dtype=np.int32)
    if tol is not None:
        orc.vacuum_assign(rho, v0, tol, rho, dV)
    mx, ref_on = orc.bader_calc('ongrid', rho, v0, dm, T)
    ref_ng = ref_on.astype(np.int32)
    log = []
    orc.refine('neargrid', ('all', -1), rho, ref_ng, dm, T, log=log)
    parts = [np.load(os.path.join(str(tmp_path), f'r{r}.npz')) for r in range(world)]
    lab_on = np.concatenate([p['lab_on'] for p in parts], axis=0)
    lab_ng = np.concatenate([p['lab_ng'] for p in parts], axis=0)
    assert [int(p['x0']) for p in parts] == sorted(int(p['x0']) for p in parts)
    for p in parts:
        np.testing.assert_array_equal(p['maxima'], mx)          # same list on every rank
    np.testing.assert_array_equal(lab_on, ref_on)                # ongrid: bit-exact, same numbering
    np.testing.assert_array_equal(lab_ng, ref_ng)                # Jacobi passes: partition independent
    assert [tuple(h) for h in parts[0]['hist']][:len(log)] == log[:len(parts[0]['hist'])]
    # 'changed' mode: edge_check's centre selection follows the global scan order
    ref_c = ref_on.astype(np.int32)
    log_c = []
    orc.refine('neargrid', ('changed', 3), rho, ref_c, dm, T, log=log_c)
    lab_c = np.concatenate([p['lab_c'] for p in parts], axis=0)
    np.testing.assert_array_equal(lab_c, ref_c)
    assert [tuple(h) for h in parts[0]['hist_c']] == log_c, (parts[0]['hist_c'], log_c)
    n = mx.shape[0]
    q, v = np.zeros(n), np.zeros(n)
    orc.charge_sum(q, v, dV, rho, ref_ng)
    np.testing.assert_allclose(parts[0]['q'], q, rtol=1e-12)
    np.testing.assert_allclose(parts[0]['v'], v, rtol=1e-12)


def test_slab_bounds():
    from pybader_b200.sharded import slab_bounds
    assert slab_bounds(10, 3) == [0, 4, 7, 10]
    assert slab_bounds(2048, 8) == [256 * i for i in range(9)]
