"""The reference-shaped entry points over several GPUs (WORLD_SIZE > 1).

`pybader_b200.thread_handlers` / `pybader_b200.utils` dispatch here when the process runs
under `torchrun` (one process per GPU).  The program is SPMD like every torch.distributed
program: each rank runs the same unmodified `Bader` object on the same host arrays, uploads
only its own x-slab window (plus halo planes), takes part in the sharded analysis
(`pybader_b200.sharded.ShardedBader`: slab seed, exit resolution and numbering over NCCL,
round loops inside the library) and receives the complete label volume back through an
all-gather -- so `bader_volumes`, `atoms_volumes`, the sums and the distances are identical
on every rank and identical to a one-GPU run (the 2/4/8-rank tests compare them bit for bit).

    thread_handlers.bader_calc / refine / assign_to_atoms / surface_distance
                                              thread_handlers.py:15-75, 128-236, 78-125, 239-297
    utils.vacuum_assign / charge_sum / volume_mask          utils.py:383-401, 236-252, 462-476

Grids that do not fit one host (2048^3) use `ShardedBader` directly with device-generated or
per-slab inputs, as bench.py does.
"""
import ctypes
import os

import numpy as np

from . import session as _single
from .engine import LABELS_ATOMS, LABELS_BADER, METHODS, REFINE_METHODS, RHO_CHARGE, RHO_REFERENCE, RHO_SPIN

_state = {}
DEFAULT_HALO = 4


def world_size():
    return int(os.environ.get('WORLD_SIZE', '1'))


def active():
    """True when the handlers should run sharded: launched by torchrun with more than one rank"""
    return world_size() > 1 and not os.environ.get('BDR_FORCE_SINGLE')


def _dist():
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return dist


class ShardedSession:
    """device residency of one grid shape across the ranks (the sharded counterpart of
    pybader_b200.session.Session: same full-content keys, per-rank windows on the device)"""

    def __init__(self, shape, halo=DEFAULT_HALO):
        import torch
        from .sharded import Comm, ShardedBader, SlabBackend
        _dist()
        local = int(os.environ.get('LOCAL_RANK', '0'))
        self.shape = tuple(int(s) for s in shape)
        self.sb = ShardedBader(self.shape, Comm(), lambda ws, h: SlabBackend(ws, h, device=local), halo=halo)
        self.be = self.sb.backend
        self.torch = torch
        self.rho_key = [None, None, None]
        self.label_key = [None, None]
        self.n_max = 0

    def close(self):
        self.be.close()

    # ---- host -> window ------------------------------------------------------
    def window(self, arr):
        return np.ascontiguousarray(np.asarray(arr)[self.sb.window_x])

    def density_slot(self, arr, prefer):
        key = _single.fingerprint(arr)
        for slot in (RHO_REFERENCE, RHO_CHARGE, RHO_SPIN):
            if self.rho_key[slot] == key:
                return slot
        w = np.ascontiguousarray(self.window(arr), dtype=np.float64)
        self.be.check(self.be.lib.bdr_upload_density(self.be.h, prefer, w.ctypes.data))
        self.rho_key[prefer] = key
        return prefer

    def reference(self, arr):
        """`arr` becomes the reference density (slot RHO_REFERENCE, the only slot the maximum
        search, the refinement, the vacuum mask and the surface distance read)"""
        key = _single.fingerprint(arr)
        if self.rho_key[RHO_REFERENCE] == key:
            return RHO_REFERENCE
        for slot in (RHO_CHARGE, RHO_SPIN):
            if self.rho_key[slot] == key:
                self.be.check(self.be.lib.bdr_copy_density(self.be.h, RHO_REFERENCE, slot))
                self.rho_key[RHO_REFERENCE] = key
                return RHO_REFERENCE
        self.rho_key[RHO_REFERENCE] = None
        return self.density_slot(arr, RHO_REFERENCE)

    def free_density_slot(self):
        for slot in (RHO_REFERENCE, RHO_CHARGE, RHO_SPIN):
            if self.rho_key[slot] is None:
                return slot
        return RHO_SPIN

    def labels_in(self, arr, which):
        """make label set `which` of every window mirror the host array `arr`"""
        arr = np.asarray(arr)
        digest, all_zero = _single.content_hash(arr)
        key = (arr.shape, arr.dtype.str, digest)
        if self.label_key[which] == key:
            return
        if all_zero:
            self.be.check(self.be.lib.bdr_clear_labels(self.be.h, which))
            self.be.check(self.be.lib.bdr_synchronize(self.be.h))
        else:
            w = self.window(arr)
            if w.dtype.kind == 'u':
                w = w.astype(np.int64)
            self.be.check(self.be.lib.bdr_upload_labels(self.be.h, which, w.ctypes.data, w.dtype.itemsize))
        self.label_key[which] = key

    def label_slot(self, arr):
        """slot that already mirrors `arr`, else BADER after an upload"""
        key = _single.fingerprint(arr)
        for slot in (LABELS_BADER, LABELS_ATOMS):
            if self.label_key[slot] == key:
                return slot
        self.labels_in(arr, LABELS_BADER)
        return LABELS_BADER

    # ---- window -> host --------------------------------------------------------
    def _owned(self, which):
        be, sb = self.be, self.sb
        from .sharded import _DevArray
        t = self.torch.as_tensor(_DevArray(be._ptr(3 + which), be.shape, '<i4'), device=be.device)
        return t[sb.halo:sb.halo + sb.nxl]

    def gather(self, owned, np_dtype):
        """all-gather of the ranks' owned planes [nxl_r, ny, nz] into the full host volume"""
        torch, sb = self.torch, self.sb
        dist = _dist()
        tdt = {'int8': torch.int8, 'int16': torch.int16, 'int32': torch.int32, 'int64': torch.int64,
               'float64': torch.float64}[np.dtype(np_dtype).name]
        mine = owned.to(tdt).contiguous()
        nmax = max(sb.bounds[r + 1] - sb.bounds[r] for r in range(sb.comm.world))
        if mine.shape[0] < nmax:
            pad = torch.zeros((nmax - mine.shape[0],) + tuple(mine.shape[1:]), dtype=tdt, device=mine.device)
            mine = torch.cat([mine, pad])
        parts = [torch.empty_like(mine) for _ in range(sb.comm.world)]
        dist.all_gather(parts, mine, group=sb.comm.group)
        out = np.empty(self.shape, dtype=np_dtype)
        for r, p in enumerate(parts):
            n = sb.bounds[r + 1] - sb.bounds[r]
            out[sb.bounds[r]:sb.bounds[r + 1]] = p[:n].cpu().numpy()
        return out

    def labels_out(self, which, np_dtype, out=None):
        host = self.gather(self._owned(which), np.dtype(np_dtype) if out is None else out.dtype)
        if out is not None:
            out[...] = host
            host = out
        self.label_key[which] = _single.fingerprint(host)
        return host

    def allreduce(self, arr, op='sum'):
        dist = _dist()
        t = self.torch.as_tensor(np.ascontiguousarray(arr), device=self.be.device)
        dist.all_reduce(t, op={'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN,
                               'max': dist.ReduceOp.MAX}[op], group=self.sb.comm.group)
        return t.cpu().numpy()


def get(shape):
    shape = tuple(int(s) for s in shape)
    s = _state.get(shape)
    if s is None:
        close_all()
        s = ShardedSession(shape, int(os.environ.get('BDR_HALO', DEFAULT_HALO)))
        _state[shape] = s
    return s


def close_all():
    for s in list(_state.values()):
        s.close()
    _state.clear()


# ---- the entry points ------------------------------------------------------------
def vacuum_assign(reference, volumes, vac_tol, density, voxel_volume):
    s = get(reference.shape)
    s.reference(reference)
    dslot = s.density_slot(density, RHO_CHARGE)
    s.labels_in(volumes, LABELS_BADER)
    q, v = ctypes.c_double(0), ctypes.c_double(0)
    s.be.check(s.be.lib.bdr_vacuum_assign(s.be.h, float(vac_tol), float(voxel_volume), dslot,
                                          ctypes.byref(q), ctypes.byref(v)))
    tot = s.allreduce(np.array([q.value, v.value]))
    if tot[1] != 0.0:
        s.labels_out(LABELS_BADER, volumes.dtype, out=volumes)
    return volumes, float(tot[0]), float(tot[1])


def bader_calc(method, density, volumes, dist_mat, T_grad, threads=1):
    from .utils import dtype_calc
    if method not in METHODS:
        raise AttributeError(f"module 'pybader.methods' has no attribute '{method}'")
    s = get(density.shape)
    s.reference(density)
    s.labels_in(volumes, LABELS_BADER)
    if method == 'ongrid':
        mx = s.sb.ongrid(dist_mat)
    else:
        mx = s.sb.neargrid(dist_mat, T_grad)
        if not s.sb.settled:
            raise RuntimeError("bader_calc(neargrid): the sharded rounds did not settle")
    s.n_max = int(mx.shape[0])
    out = s.labels_out(LABELS_BADER, dtype_calc(-max(s.n_max, 0)))
    return np.asarray(mx, dtype=np.int64), out


def refine(method, refine_mode, density, volumes, dist_mat, T_grad, threads=1):
    if method not in REFINE_METHODS:
        return
    check_mode, iters = tuple(refine_mode)
    if iters == 0:
        return
    mode = 'all' if check_mode.lower() == 'all' else 'changed'
    s = get(density.shape)
    s.reference(density)
    slot = s.label_slot(volumes)
    if slot != LABELS_BADER:
        # the sharded loops work on the BADER label set: move the atom labels there
        s.labels_in(volumes, LABELS_BADER)
    history = s.sb.refine(dist_mat, T_grad, iters, mode=mode)
    if history and any(ch for _, ch in history):
        s.labels_out(LABELS_BADER, volumes.dtype, out=volumes)
    refine.last_history = history


refine.last_history = []


def assign_to_atoms(bader_max, atoms, lattice, volumes, threads=1):
    from .utils import dtype_calc
    s = get(volumes.shape)
    s.labels_in(volumes, LABELS_BADER)
    m = np.ascontiguousarray(bader_max, dtype=np.float64).reshape(-1, 3)
    a = np.ascontiguousarray(atoms, dtype=np.float64).reshape(-1, 3)
    lat = np.ascontiguousarray(lattice, dtype=np.float64)
    who = np.zeros(m.shape[0], dtype=np.int64)
    dist_ = np.zeros(m.shape[0], dtype=np.float64)
    s.be.check(s.be.lib.bdr_assign_atoms(s.be.h, m.ctypes.data, m.shape[0], a.ctypes.data, a.shape[0],
                                         lat.ctypes.data, who.ctypes.data, dist_.ctypes.data))
    atoms_volumes = s.labels_out(LABELS_ATOMS, dtype_calc(-a.shape[0]))
    return who, dist_, atoms_volumes


def surface_distance(density, volumes, lattice, atoms, threads=1):
    s = get(density.shape)
    s.reference(density)
    key = _single.fingerprint(volumes)
    which = LABELS_ATOMS if s.label_key[LABELS_ATOMS] == key else None
    if which is None:
        which = LABELS_BADER
        s.labels_in(volumes, LABELS_BADER)
    a = np.ascontiguousarray(atoms, dtype=np.float64).reshape(-1, 3)
    lat = np.ascontiguousarray(lattice, dtype=np.float64)
    best = np.zeros(a.shape[0], dtype=np.float64)
    seen = np.zeros(a.shape[0], dtype=np.int64)
    edges = ctypes.c_int64(0)
    s.be.check(s.be.lib.bdr_slab_surface_distance(s.be.h, which, lat.ctypes.data, a.ctypes.data, a.shape[0],
                                                  best.ctypes.data, seen.ctypes.data, ctypes.byref(edges)))
    best = s.allreduce(best, 'min')
    seen = s.allreduce(seen, 'max')
    total = s.allreduce(np.array([edges.value], dtype=np.int64))[0]
    if total == 0:
        return None
    return np.where(seen > 0, np.sqrt(best), 0.0)


def charge_sum(charge, volume, voxel_volume, density, volumes):
    s = get(volumes.shape)
    lslot = s.label_slot(volumes)
    dslot = s.density_slot(density, s.free_density_slot())
    n = charge.shape[0]
    q, v = np.zeros(n), np.zeros(n)
    s.be.check(s.be.lib.bdr_charge_sum(s.be.h, lslot, dslot, float(voxel_volume), n, q.ctypes.data,
                                       v.ctypes.data))
    tot = s.allreduce(np.stack([q, v]))
    # the reference adds into the caller's (zeroed) arrays and scales the charge afterwards
    # (utils.py:246-252); the per-rank sums above are already scaled
    charge[...] = charge * voxel_volume + tot[0]
    volume[...] = volume + tot[1]


def volume_mask(volumes, density, vol_num):
    s = get(volumes.shape)
    lslot = s.label_slot(volumes)
    dslot = s.density_slot(density, s.free_density_slot())
    w = np.empty(s.be.shape, dtype=np.float64)
    s.be.check(s.be.lib.bdr_volume_mask(s.be.h, lslot, dslot, int(vol_num), w.ctypes.data))
    owned = s.torch.as_tensor(w[s.sb.halo:s.sb.halo + s.sb.nxl], device=s.be.device)
    return s.gather(owned, np.float64)
