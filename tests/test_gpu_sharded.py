"""Sharded path against the single-GPU path on 2, 4 and 8 GPUs (SURVEY.md 8e "Check":
P-GPU labels bit-identical to 1-GPU labels at sizes one GPU holds).  Each case needs that
many visible GPUs and is skipped otherwise: run with `gpurun --gpus N`."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = {
    # name: (world, shape, halo); *_pyproto drives the round loops from Python
    # (BDR_PY_PROTOCOL=1, the code path the gloo CPU tests cover) instead of from the library
    'w2_toy_halo8': (2, (96, 64, 80), 8),
    'w2_toy_halo8_pyproto': (2, (96, 64, 80), 8),
    'w2_512_halo4': (2, (512, 256, 256), 4),      # the bench's halo
    'w4_512_halo4': (4, (512, 256, 256), 4),
    'w8_512_halo4': (8, (512, 256, 256), 4),
}


def _case(shape):
    from pybader_b200 import geometry as geo, synth
    if shape == (96, 64, 80):
        c = synth.case_rocksalt(96, cells=2, offset=0.13, a=5.64)
        c['shape'] = shape
        rho, atoms = synth.make(c)
    else:
        # jittered lattice sites in an orthorhombic cell: separable, generated from tables
        c = synth.case_lattice_sites(shape, (4, 2, 2), (20.0, 10.0, 10.0), seed=11)
        tx, ty, tz = synth.separable_tables(c)
        rho = np.zeros(shape)
        for a in range(tx.shape[0]):
            rho += (tx[a][:, None, None] * ty[a][None, :, None]) * tz[a][None, None, :]
    return rho, geo.distance_matrix(c['lattice'], rho.shape), geo.T_grad(c['lattice'], rho.shape), \
        geo.voxel_volume(c['lattice'], rho.shape)


def _worker(rank, world, port, out, shape, halo):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        from pybader_b200.sharded import Comm, ShardedBader, SlabBackend
        rho, dm, T, dV = _case(shape)
        sb = ShardedBader(rho.shape, Comm(), lambda ws, h: SlabBackend(ws, h, device=rank), halo=halo)
        win = np.ascontiguousarray(rho[sb.window_x])
        sb.backend.check(sb.backend.lib.bdr_upload_density(sb.backend.h, 0, win.ctypes.data))
        sb.backend.clear_labels()
        mx = sb.ongrid(dm)
        lab_on = sb.owned_labels().cpu().numpy().copy()
        hist = sb.refine(dm, T, -1)
        lab_ng = sb.owned_labels().cpu().numpy().copy()
        q, v = sb.charge_sum(mx.shape[0], dV)
        # 'changed'-mode refinement across the slabs (edge_check's centre selection follows the
        # global scan order): from the same ongrid labels
        sb.backend.clear_labels()
        sb.ongrid(dm)
        hist_c = sb.refine(dm, T, 3, mode='changed')
        lab_c = sb.owned_labels().cpu().numpy().copy()
        # bader_calc('neargrid') of the sharded path: incremental rounds, then the caller's
        # refine(('changed', 2)) like the bench step
        sb.backend.clear_labels()
        mx2 = sb.neargrid(dm, T)
        assert sb.settled
        hist2 = sb.refine(dm, T, 2, mode='changed')
        lab_nn = sb.owned_labels().cpu().numpy().copy()
        q2, v2 = sb.charge_sum(mx2.shape[0], dV)
        np.savez(os.path.join(out, f'r{rank}.npz'), maxima=mx, lab_on=lab_on, lab_ng=lab_ng,
                 hist=np.array(hist), q=q, v=v, maxima2=mx2, lab_nn=lab_nn, hist2=np.array(hist2),
                 q2=q2, v2=v2, lab_c=lab_c, hist_c=np.array(hist_c))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', list(CASES))
def test_sharded_labels_equal_single_gpu(tmp_path, name):
    import torch
    import torch.multiprocessing as mp
    world, shape, halo = CASES[name]
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from pybader_b200 import build
    build.build()
    if name.endswith('_pyproto'):
        os.environ['BDR_PY_PROTOCOL'] = '1'
    else:
        os.environ.pop('BDR_PY_PROTOCOL', None)
    try:
        mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), shape, halo), nprocs=world, join=True)
    finally:
        os.environ.pop('BDR_PY_PROTOCOL', None)
    from pybader_b200.engine import Engine, LABELS_BADER
    rho, dm, T, dV = _case(shape)
    e = Engine(rho.shape)
    e.upload_density(0, rho)
    e.clear_labels()
    mx = e.bader_calc('ongrid', dm, T)
    ref_on = e.download_labels(LABELS_BADER, np.int32)
    hist = e.refine(LABELS_BADER, 'all', -1, dm, T)
    ref_ng = e.download_labels(LABELS_BADER, np.int32)
    q, v = np.zeros(mx.shape[0]), np.zeros(mx.shape[0])
    e.charge_sum(LABELS_BADER, 0, dV, q, v)
    e.upload_labels(LABELS_BADER, ref_on)
    hist_c = e.refine(LABELS_BADER, 'changed', 3, dm, T)
    ref_c = e.download_labels(LABELS_BADER, np.int32)
    e.clear_labels()
    mxn = e.bader_calc('neargrid', dm, T)
    hist_n = e.refine(LABELS_BADER, 'changed', 2, dm, T)
    ref_nn = e.download_labels(LABELS_BADER, np.int32)
    qn, vn = np.zeros(mxn.shape[0]), np.zeros(mxn.shape[0])
    e.charge_sum(LABELS_BADER, 0, dV, qn, vn)
    e.close()
    parts = [np.load(os.path.join(str(tmp_path), f'r{r}.npz')) for r in range(world)]
    np.testing.assert_array_equal(parts[0]['maxima'], mx)
    np.testing.assert_array_equal(np.concatenate([p['lab_on'] for p in parts]), ref_on)
    np.testing.assert_array_equal(np.concatenate([p['lab_ng'] for p in parts]), ref_ng)
    assert [tuple(h) for h in parts[0]['hist']] == hist
    np.testing.assert_allclose(parts[0]['q'], q, rtol=1e-12)
    np.testing.assert_allclose(parts[0]['v'], v, rtol=1e-12)
    np.testing.assert_array_equal(np.concatenate([p['lab_c'] for p in parts]), ref_c)
    assert [tuple(h) for h in parts[0]['hist_c']] == hist_c, (parts[0]['hist_c'], hist_c)
    # sharded bader_calc('neargrid') + one exact pass vs the single-GPU one: same
    # maxima, the exact pass finds (next to) nothing to do, labels >= 99.9 % equal
    np.testing.assert_array_equal(parts[0]['maxima2'], mxn)
    assert int(parts[0]['hist2'][0][1]) <= max(2, 1e-4 * ref_nn.size)
    lab_nn = np.concatenate([p['lab_nn'] for p in parts])
    assert np.mean(lab_nn == ref_nn) >= 0.999
    print(f"{name}: ongrid, ongrid+refine(all,-1) ({len(hist)} passes) and ongrid+refine(changed,3) "
          f"(history {hist_c}) labels bit-identical to 1 GPU over {world} ranks; neargrid+refine(changed,2): "
          f"{int(np.count_nonzero(lab_nn != ref_nn))} of {ref_nn.size} voxels differ, history "
          f"{[tuple(h) for h in parts[0]['hist2']]} vs {hist_n} on 1 GPU")
    np.testing.assert_allclose(parts[0]['q2'], qn, rtol=1e-6)
    np.testing.assert_allclose(parts[0]['v2'], vn, rtol=1e-6)


# ---- the reference-shaped handlers over two ranks: unmodified Bader.__call__ under "torchrun" ----
def _bader_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    import pybader_b200
    from baseline.refload import import_reference
    from pybader_b200 import sharded_handlers
    import tests.test_gpu_boundary as tb
    ref = import_reference(scratch=f'/tmp/pybader_ref_home_r{rank}')
    old = pybader_b200.install(ref['interface'], readers=False)
    try:
        assert sharded_handlers.active()
        # BASELINE config 1, DEFAULT profile
        g = tb.load('c1_default')
        rho = tb.density_from_tables(g['tx'], g['ty'], g['tz'])
        b = tb.make_bader(ref, {'charge': rho}, g)
        tb.compare(b, g, labels_exact=False)
        d1 = int(np.count_nonzero(b.bader_volumes != g['bader_volumes']))
        # utils.volume_mask over the ranks (the export path, interface.py:600-621)
        from pybader_b200 import utils as ut
        m = ut.volume_mask(b.bader_volumes, rho, 1)
        np.testing.assert_array_equal(m, np.where(b.bader_volumes == 1, rho, 0.0))
        # the `speed` profile: ongrid + refine ('changed', 3) on atoms_volumes, bit-exact
        g2 = tb.load('c1_speed')
        b2 = tb.make_bader(ref, {'charge': rho}, g2, profile='speed')
        tb.compare(b2, g2, labels_exact=True)
        # reference is not density, vacuum_tol, spin
        g3 = tb.load('ref_spin')
        b3 = tb.make_bader(ref, {'charge': g3['charge'].copy(), 'spin': g3['spin'].copy()}, g3,
                           reference=g3['reference'].copy(), vacuum_tol=float(g3['vacuum_tol']), spin_flag=True)
        tb.compare(b3, g3, labels_exact=False)
        np.savez(os.path.join(out, f'b{rank}.npz'), differ=d1, atoms_charge=b.atoms_charge,
                 surface=b.atoms_surface_distance, labels=b.bader_volumes)
    finally:
        pybader_b200.uninstall(old, ref['interface'])
        sharded_handlers.close_all()
        if dist.is_initialized():
            dist.destroy_process_group()


def test_unmodified_bader_call_over_two_ranks(tmp_path):
    """the sharded pipeline behind the API.  Two processes (what torchrun would
    start) run the UNMODIFIED reference `Bader.__call__` after pybader_b200.install(); the
    handlers dispatch to the sharded engine, every rank gets the complete results, and they
    match the golden outputs of the reference's own run (tests/golden_call)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from baseline.refload import find_reference
    if find_reference() is None:
        pytest.skip("reference not installed")
    from pybader_b200 import build
    build.build()
    mp.spawn(_bader_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(str(tmp_path), f'b{r}.npz')) for r in range(2)]
    np.testing.assert_array_equal(parts[0]['labels'], parts[1]['labels'])      # same on every rank
    np.testing.assert_array_equal(parts[0]['atoms_charge'], parts[1]['atoms_charge'])
    print(f"unmodified Bader.__call__ over 2 ranks: c1 DEFAULT {int(parts[0]['differ'])} of "
          f"{parts[0]['labels'].size} voxels differ from the reference; speed profile bit-exact; "
          f"reference != density + spin within the bars")
