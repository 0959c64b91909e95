"""Sharded (multi-GPU) Bader analysis: one process per GPU, x-slabs, NCCL.

The reference parallelises by cutting the volume into bricks, letting every
brick run on its own and resolving trajectories that leave a brick afterwards
(thread_handlers.py:27-75; "edge" labels methods.py:170-199; utils.edge_assign
utils.py:263-280; renumbering utils.volume_offset utils.py:497-510).  The same
idea, restated for GPUs:

* the grid is cut into contiguous slabs along the slowest storage axis (x);
  rank r owns planes [x0_r, x1_r) and keeps `halo` extra planes on each side;
* every rank runs the single-GPU kernels on its window.  The outermost plane
  on each side is an *exit*: an ascent path reaching it stops there;
* exits are resolved between neighbours by exchanging one plane of resolved
  root ids per direction, iterated until no exit is left (<= world rounds);
* volumes are numbered globally by their first voxel (C order), exactly like
  the single-GPU path, from an all-gather of (root id, first voxel) pairs;
* refinement runs full Jacobi passes; after each pass the owned boundary
  planes of the label array are sent to the neighbours' halos.

`torch.distributed` is the plumbing (process group, send/recv, all-reduce); all
per-voxel work is done by libbader_b200.so kernels through `SlabBackend`.
The protocol code below only touches plane- and root-sized tensors and is
device agnostic, so tests drive it on CPU over gloo with a model backend.
"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

UNRESOLVED = -1
VACUUM = -2
NO_VOXEL = 0x7f7f7f7f


def slab_bounds(nx, world):
    """contiguous split of nx planes; the first nx % world ranks get one more"""
    base, rem = divmod(nx, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


class Comm:
    """ring neighbours over torch.distributed (NCCL on GPUs, gloo in CPU tests)"""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.prev = (self.rank - 1) % self.world
        self.next = (self.rank + 1) % self.world

    def ring_exchange(self, send_up, send_down, recv_lo, recv_hi):
        """send_up -> next rank's recv_lo, send_down -> prev rank's recv_hi"""
        if self.world == 1:
            recv_lo.copy_(send_up)
            recv_hi.copy_(send_down)
            return
        ops = [dist.P2POp(dist.isend, send_up, self.next, self.group, tag=1),
               dist.P2POp(dist.isend, send_down, self.prev, self.group, tag=2),
               dist.P2POp(dist.irecv, recv_lo, self.prev, self.group, tag=1),
               dist.P2POp(dist.irecv, recv_hi, self.next, self.group, tag=2)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def allreduce_sum(self, value, device):
        t = torch.tensor([int(value)], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def allreduce_max(self, value, device):
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def allgather_padded(self, t, pad_value):
        """all-gather of 1-D tensors of different lengths (padded to the max)"""
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)
        m = int(n.item())
        buf = torch.full((m,), pad_value, dtype=t.dtype, device=t.device)
        buf[:t.shape[0]] = t
        out = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(out, buf, group=self.group)
        return torch.cat(out)


class ShardedBader:
    """The cross-rank protocol.  `backend` does the per-voxel work on this
    rank's window and exposes window-shaped torch tensors (see SlabBackend)."""

    def __init__(self, global_shape, comm, backend_factory, halo=8):
        self.comm = comm
        self.shape = tuple(int(s) for s in global_shape)
        nx, self.ny, self.nz = self.shape
        self.halo = int(halo)
        b = slab_bounds(nx, comm.world)
        self.x0, self.x1 = b[comm.rank], b[comm.rank + 1]
        self.nxl = self.x1 - self.x0
        if min(b[i + 1] - b[i] for i in range(comm.world)) < self.halo:
            raise ValueError("slabs thinner than the halo: use fewer ranks or a smaller halo")
        self.W = self.nxl + 2 * self.halo
        self.plane = self.ny * self.nz
        # global x of every window plane (periodic)
        self.window_x = (np.arange(self.W) + self.x0 - self.halo) % nx
        self.backend = backend_factory((self.W, self.ny, self.nz), self.halo)
        self.maxima = np.zeros((0, 3), dtype=np.int64)
        self.bounds = b
        # trajectories that leave this rank's window continue on the owner's
        # memory (CUDA IPC over NVLink); the CPU model backend has no such need
        if hasattr(self.backend, 'ipc_export'):
            self._attach_peers()
        # the library's own NCCL communicator: the round loops of neargrid() and refine() then
        # run inside libbader_b200.so (bdr_slab_rounds / bdr_slab_refine) with one host
        # decision per round; BDR_PY_PROTOCOL=1 keeps the loops below (what the CPU tests drive)
        self.native_loops = False
        if hasattr(self.backend, 'comm_init') and not os.environ.get('BDR_PY_PROTOCOL'):
            self._init_native_comm()

    def _attach_peers(self):
        be = self.backend
        mine = torch.frombuffer(bytearray(be.ipc_export()), dtype=torch.uint8).to(be.device)
        if self.comm.world > 1:
            allh = [torch.empty_like(mine) for _ in range(self.comm.world)]
            dist.all_gather(allh, mine, group=self.comm.group)
            allh = torch.cat(allh)
        else:
            allh = mine
        be.ipc_attach(self.comm.world, self.comm.rank, bytes(allh.cpu().numpy().tobytes()),
                      self.bounds, self.shape[0])

    def _init_native_comm(self):
        be = self.backend
        ident = torch.zeros(128, dtype=torch.uint8, device=be.device)
        if self.comm.rank == 0:
            ident.copy_(torch.frombuffer(bytearray(be.comm_id()), dtype=torch.uint8))
        if self.comm.world > 1:
            dist.broadcast(ident, src=0, group=self.comm.group)
        be.comm_init(self.comm.world, self.comm.rank, bytes(ident.cpu().numpy().tobytes()))
        self.native_loops = True

    # ---- helpers -----------------------------------------------------------
    def _gid_of_window_index(self, widx):
        """window-linear voxel index -> global linear voxel index (int64)"""
        xw = torch.div(widx, self.plane, rounding_mode='floor')
        rest = widx - xw * self.plane
        wx = torch.as_tensor(self.window_x, dtype=torch.int64, device=widx.device)
        return wx[xw] * self.plane + rest

    def exchange_halo(self, t):
        """fill the halo planes of a window tensor [W, ny, nz] from the owners"""
        H, n = self.halo, self.nxl
        up = t[n:n + H].contiguous()        # my top owned planes -> next's low halo
        down = t[H:2 * H].contiguous()      # my bottom owned planes -> prev's high halo
        lo = torch.empty_like(up)
        hi = torch.empty_like(down)
        self.comm.ring_exchange(up, down, lo, hi)
        t[:H].copy_(lo)
        t[self.W - H:].copy_(hi)

    def _phase(self, name):
        """BDR_PHASES=1: wall-clock per phase of the protocol (device drained at every
        boundary, so the step runs slower while this is on)"""
        if not os.environ.get('BDR_PHASES'):
            return
        import time
        torch.cuda.synchronize()
        if hasattr(self.backend, '_sync'):
            self.backend.check(self.backend.lib.bdr_synchronize(self.backend.h))
        now = time.perf_counter()
        acc = self.__dict__.setdefault('phase_ms', {})
        if getattr(self, '_phase_name', None) is not None:
            acc[self._phase_name] = acc.get(self._phase_name, 0.0) + (now - self._phase_t0) * 1e3
        self._phase_name, self._phase_t0 = name, now

    # ---- ongrid: seed, exits, numbering -----------------------------------
    def _dbg(self, msg):
        if os.environ.get('BDR_DEBUG'):
            print(f"[sharded r{self.comm.rank}] {msg}", file=sys.stderr, flush=True)

    def ongrid(self, dist_mat, method='ongrid', provisional=False):
        """seed + exit resolution + global numbering.  Volumes are numbered by their first
        voxel in C order (the reference's discovery order).  With `provisional` they are
        numbered by the voxel index of their maximum instead, without the first-voxel pass:
        bader_calc('neargrid') numbers once its labels are final (`renumber`), like the
        single-GPU path does."""
        be, H, P = self.backend, self.halo, self.plane
        self._dbg("seed")
        self._phase('seed')
        n_real, exit_base = be.seed(dist_mat, method)
        self._phase('exits')
        self._dbg(f"seeded: {n_real} local maxima")
        codes = be.labels()                      # int32 [W, ny, nz]: -1 vacuum, -2-s slots
        dev = codes.device
        n_slots = exit_base + n_real
        # slot -> global root id
        G = torch.full((n_slots,), UNRESOLVED, dtype=torch.int64, device=dev)
        if n_real:
            G[exit_base:] = self._gid_of_window_index(be.roots().to(torch.int64))

        vac = torch.tensor(VACUUM, dtype=torch.int64, device=dev)

        def export(plane_index):
            c = codes[plane_index].reshape(-1).to(torch.int64)
            # (no boolean-mask indexing here: it would synchronise the host on every call)
            return torch.where(c <= -2, G[(-2 - c).clamp_(min=0)], vac)

        # one round per slab boundary an ascent path crosses (paths are acyclic, so
        # this ends; a path winding along a boundary can need more than `world`)
        self.exit_rounds = 0
        for _ in range(64 + 2 * self.comm.world):
            self.exit_rounds += 1
            up, down = export(self.nxl), export(2 * H - 1)
            lo, hi = torch.empty_like(up), torch.empty_like(down)
            self.comm.ring_exchange(up, down, lo, hi)
            G[:P] = lo
            G[P:2 * P] = hi
            pending = ((up == UNRESOLVED).sum() + (down == UNRESOLVED).sum()).reshape(1)
            if self.comm.world > 1:
                dist.all_reduce(pending, op=dist.ReduceOp.SUM, group=self.comm.group)
            pending = int(pending.item())
            if os.environ.get('BDR_DEBUG') and self.comm.rank == 0:
                print(f"[sharded] exit round {self.exit_rounds}: pending {pending}", file=sys.stderr,
                      flush=True)
            if pending == 0:
                break
        else:
            raise RuntimeError("exit resolution did not converge")
        # one more hop is implied: exits used by owned voxels are now resolved,
        # because every plane a neighbour needs from me was exported resolved

        # first owned voxel of every slot -> per root id
        self._dbg("numbering")
        self._phase('numbering')
        if provisional:
            gids = self.comm.allgather_padded(torch.unique(G[G >= 0]), -1)
            uniq = torch.unique(gids[gids >= 0])
            order = torch.arange(uniq.shape[0], device=dev)
        else:
            first_w = be.first_voxel(n_slots).to(torch.int64)          # window-linear or NO_VOXEL
            used = first_w != NO_VOXEL
            if bool((G[used] == UNRESOLVED).any()):
                raise RuntimeError("an owned voxel ends in an unresolved exit")
            sel = used & (G >= 0)
            gid_l = G[sel]
            first_l = self._gid_of_window_index(first_w[sel])
            gids = self.comm.allgather_padded(gid_l, -1)
            firsts = self.comm.allgather_padded(first_l, -1)
            keep = gids >= 0
            gids, firsts = gids[keep], firsts[keep]
            uniq, inv = torch.unique(gids, return_inverse=True)
            first_u = torch.full((uniq.shape[0],), torch.iinfo(torch.int64).max, dtype=torch.int64,
                                 device=dev)
            first_u.scatter_reduce_(0, inv, firsts, reduce='amin')
            order = torch.argsort(first_u)              # volume number -> index into uniq
        number_of = torch.empty_like(order)
        number_of[order] = torch.arange(order.shape[0], device=dev)
        # slot -> volume number (vacuum and unused exits -> -1)
        rank_lut = torch.full((max(n_slots, 1),), -1, dtype=torch.int32, device=dev)
        ok = G >= 0
        if uniq.shape[0]:
            pos = torch.searchsorted(uniq, G[ok])
            pos = pos.clamp(max=uniq.shape[0] - 1)
            hit = uniq[pos] == G[ok]
            vals = torch.where(hit, number_of[pos], torch.full_like(pos, -1)).to(torch.int32)
            rank_lut[ok] = vals
        self._dbg("apply rank")
        self._phase('apply_rank')
        be.apply_rank(rank_lut)
        self._phase(None)
        self._dbg("ongrid done")
        mg = uniq[order].cpu().numpy()
        self.maxima = np.stack([mg // P, (mg // self.nz) % self.ny, mg % self.nz], axis=1)
        return self.maxima

    def renumber(self):
        """number the volumes by the first voxel (C order) that carries them in the CURRENT
        labels and reorder `maxima` accordingly (utils.volume_offset utils.py:497-510; the
        reference's discovery order)"""
        be = self.backend
        n = int(self.maxima.shape[0])
        if n == 0 or not hasattr(be, 'first_voxel_labels'):
            return self.maxima
        dev = be.labels().device
        first_w = be.first_voxel_labels(n).to(torch.int64)
        big = torch.iinfo(torch.int64).max
        first_g = torch.where(first_w != NO_VOXEL, self._gid_of_window_index(first_w.clamp(max=be.N - 1)),
                              torch.full_like(first_w, big))
        if self.comm.world > 1:
            dist.all_reduce(first_g, op=dist.ReduceOp.MIN, group=self.comm.group)
        order = torch.argsort(first_g, stable=True)      # new number -> old number
        lut = torch.empty(n, dtype=torch.int32, device=dev)
        lut[order] = torch.arange(n, dtype=torch.int32, device=dev)
        self._renumbered = not bool((order == torch.arange(n, device=dev)).all())
        if self._renumbered:
            be.relabel(lut)
            self.maxima = self.maxima[order.cpu().numpy()]
        return self.maxima

    def renumber_changed(self):
        self.renumber()
        return getattr(self, '_renumbered', False)

    # ---- refinement ---------------------------------------------------------
    def refine(self, dist_mat, T_grad, iters=-1, mode='all'):
        """thread_handlers.refine (thread_handlers.py:128-236) across the slabs; returns
        [(edges, changed)] with global counts, like the single-GPU history.

        'all': a fresh edge pass before every Jacobi trace, until nothing changes anywhere or
        `iters` passes.  'changed': after the first pass only the neighbourhoods of the
        voxels that changed are re-classified (refinement.edge_check) -- the centre selection
        of that function follows the global scan order, so the ranks iterate it together,
        exchanging the halo planes of the known array until no voxel is undecided."""
        be, dev = self.backend, self.backend.labels().device
        changed_mode = mode.lower() != 'all' and hasattr(be, 'ec_begin')
        history = []
        if iters == 0:
            return history
        if self.native_loops:
            self._phase('refine:native')
            history = be.refine_native('changed' if changed_mode else 'all', iters, dist_mat, T_grad)
            self._phase(None)
            return history
        dbg = os.environ.get('BDR_DEBUG') and self.comm.rank == 0

        def counts(*vals):
            t = torch.tensor([int(v) for v in vals], dtype=torch.int64, device=dev)
            if self.comm.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.comm.group)
            return [int(v) for v in t.tolist()]

        it = 0
        edges = changed = 0
        while iters < 0 or it < iters:
            self._phase('refine:halo')
            self.exchange_halo(be.labels())
            if it == 0 or not changed_mode:
                self._phase('refine:edge_pass')
                # every rank's classification must be complete before any rank's
                # trajectories may read it (remote reads in the trace kernel): the
                # all-reduce of the edge count is that barrier
                edges, = counts(be.edge_pass())
                if edges == 0:
                    break
            else:
                self._phase('refine:edge_check')
                be.ec_begin()
                for _ in range(1 << 20):
                    self.exchange_halo(be.known())
                    undecided, = counts(be.ec_round())
                    if undecided == 0:
                        break
                self.exchange_halo(be.known())
                edges, = counts(be.ec_finish())
            self._phase('refine:trace')
            ch, esc = be.trace_pass(dist_mat, T_grad, want_list=changed_mode)
            self._phase('refine:reduce')
            changed, escaped = counts(ch, esc)
            if escaped:
                raise RuntimeError("a trajectory left the slab halo: raise `halo`")
            history.append((edges, changed))
            if dbg:
                print(f"[sharded] pass {it}: edges {edges} changed {changed}", file=sys.stderr, flush=True)
            it += 1
            if changed == 0:
                if it == 1 and (iters < 0 or iters >= 2):
                    # the reference runs its second iteration on the unchanged labels: a fresh
                    # edge_find finds the same edges ('all'), edge_check nothing ('changed')
                    history.append((0, 0) if changed_mode else (edges, 0))
                break
        self._phase('refine:halo')
        self.exchange_halo(be.labels())
        self._phase(None)
        return history

    def neargrid(self, dist_mat, T_grad, max_passes=64):
        """ongrid seed, then Jacobi rounds until nothing changes (at most
        `max_passes`; `self.settled` says whether the last one was quiet).
        A backend with `requeue` re-traces only the edges next to voxels that
        moved (like the single-GPU bader_calc); otherwise full passes."""
        be = self.backend
        self.ongrid(dist_mat, 'neargrid', provisional=hasattr(be, 'first_voxel_labels'))
        if not hasattr(be, 'requeue'):
            hist = self.refine(dist_mat, T_grad, max_passes)
            self.settled = not hist or hist[-1][1] == 0 or hist[-1][0] == 0
            self.neargrid_history = hist
            return self.maxima
        dev, H, P = be.labels().device, self.halo, self.plane
        lab = be.labels()
        if self.native_loops:
            self._phase('rounds:native')
            hist, self.settled = be.rounds_native(dist_mat, T_grad, max_passes)
            self._phase('rounds:number')
            self.renumber()      # the LUT pass covers the whole window: halo labels stay consistent
            self._phase(None)
            self.neargrid_history = hist
            return self.maxima
        self._phase('rounds:halo')
        self.exchange_halo(lab)
        self._phase('rounds:first_pass')
        edges = self.comm.allreduce_sum(be.first_pass(), dev)      # also the barrier before remote reads
        self._phase('rounds:trace')
        changed = self.comm.allreduce_sum(be.trace(dist_mat, T_grad, True), dev) if edges else 0
        hist = [(edges, changed)]
        self._dbg(f"first pass: edges {edges} changed {changed}")
        while changed > 0 and len(hist) < max_passes:
            # the planes next to the owned slab, before and after the exchange
            self._phase('rounds:halo')
            old_lo, old_hi = lab[H - 1].clone(), lab[self.W - H].clone()
            self.exchange_halo(lab)
            self._phase('rounds:requeue')
            lo = torch.nonzero((lab[H - 1] != old_lo).reshape(-1)).reshape(-1) + (H - 1) * P
            hi = torch.nonzero((lab[self.W - H] != old_hi).reshape(-1)).reshape(-1) + (self.W - H) * P
            extra = torch.cat([lo, hi]).to(torch.int32)
            queued = self.comm.allreduce_sum(be.requeue(extra), dev)
            self._phase('rounds:trace')
            changed = self.comm.allreduce_sum(be.trace(dist_mat, T_grad, True), dev)
            hist.append((queued, changed))
            self._dbg(f"round {len(hist) - 1}: queued {queued} changed {changed}")
        self._phase('rounds:number')
        self.renumber()
        self._phase('rounds:halo')
        self.exchange_halo(lab)
        self._phase(None)
        self.settled = changed == 0
        self.neargrid_history = hist
        return self.maxima

    def charge_sum(self, n, voxel_volume, which_density=0):
        be = self.backend
        q, v = be.charge_sum(n, voxel_volume, which_density)
        t = torch.as_tensor(np.stack([q, v]), device=be.labels().device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.comm.group)
        t = t.cpu().numpy()
        return t[0], t[1]

    def owned_labels(self):
        return self.backend.labels()[self.halo:self.halo + self.nxl]


# ---------------------------------------------------------------------------
class _DevArray:
    """zero-copy torch view of device memory owned by the C library"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


class SlabBackend:
    """per-rank kernels through the C ABI (bdr_slab_*), window tensors as
    zero-copy torch views"""

    def __init__(self, window_shape, halo, device=0):
        from . import _lib
        from ._lib import check
        self.check = check
        self.lib = _lib.load()
        self.shape = tuple(window_shape)
        self.halo = halo
        self.device = torch.device('cuda', device)
        h = ctypes.c_void_p()
        check(self.lib.bdr_slab_create(int(device), *self.shape, int(halo), ctypes.byref(h)))
        self.h = h
        # one stream for the library's kernels and torch's (NCCL-ordered) work on the window
        # tensors: no device-wide synchronisation between a halo exchange and the next kernel
        torch.cuda.set_device(self.device)
        check(self.lib.bdr_set_stream(self.h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.N = int(np.prod(self.shape))
        self.n_real = 0
        self.exit_base = 0
        self._labels = None
        self._known = None
        self._roots = None
        self._keep = []

    def close(self):
        if self.h:
            self.lib.bdr_destroy(self.h)
            self.h = None

    def ipc_export(self):
        buf = ctypes.create_string_buffer(3 * 64)
        self.check(self.lib.bdr_slab_ipc_export(self.h, buf))
        return buf.raw

    def ipc_attach(self, world, rank, all_handles, bounds, nx_global):
        b = np.ascontiguousarray(bounds, dtype=np.int64)
        self.check(self.lib.bdr_slab_ipc_attach(self.h, int(world), int(rank), all_handles,
                                                b.ctypes.data, int(nx_global)))

    def _ptr(self, what):
        p = ctypes.c_void_p()
        self.check(self.lib.bdr_device_ptr(self.h, what, ctypes.byref(p)))
        return p.value

    def density(self, which=0):
        return torch.as_tensor(_DevArray(self._ptr(which), self.shape, '<f8'), device=self.device)

    def alloc_density(self, which=0):
        """make sure slot `which` has storage and return its window view"""
        zeros = np.zeros(1)
        # a 1-atom zero table allocates the slot without a host-sized upload
        z = [np.zeros((1, n)) for n in self.shape]
        self.check(self.lib.bdr_synth_separable(self.h, which, z[0].ctypes.data, z[1].ctypes.data,
                                                z[2].ctypes.data, 1))
        del zeros
        return self.density(which)

    def synth_separable(self, which, tx, ty, tz):
        tx, ty, tz = (np.ascontiguousarray(t, dtype=np.float64) for t in (tx, ty, tz))
        self.check(self.lib.bdr_synth_separable(self.h, which, tx.ctypes.data, ty.ctypes.data,
                                                tz.ctypes.data, tx.shape[0]))

    def _sync(self):
        """the library runs on torch's current stream (bdr_set_stream): torch's copies into
        the window tensors and the library's kernels are ordered by the stream itself"""
        return

    def clear_labels(self):
        self._sync()
        self.check(self.lib.bdr_clear_labels(self.h, 0))
        self.check(self.lib.bdr_synchronize(self.h))
        self._labels = None

    def vacuum_assign(self, tol, dV):
        self._sync()
        q, v = ctypes.c_double(0), ctypes.c_double(0)
        self.check(self.lib.bdr_vacuum_assign(self.h, float(tol), float(dV), 0, ctypes.byref(q),
                                              ctypes.byref(v)))
        return q.value, v.value

    def labels(self):
        if self._labels is None:
            self._labels = torch.as_tensor(_DevArray(self._ptr(3), self.shape, '<i4'),
                                           device=self.device)
        return self._labels

    def seed(self, dist_mat, method='ongrid'):
        self._sync()
        # BDR_OPT_SLAB_SEED_METHOD: 'neargrid' takes the fp32-ranked seed stencil,
        # exactly as bdr_bader_calc does on one GPU
        self.check(self.lib.bdr_set_option(self.h, 1, 1 if method == 'neargrid' else 0))
        d = np.ascontiguousarray(dist_mat, dtype=np.float64)
        n, xb = ctypes.c_int64(0), ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_seed(self.h, d.ctypes.data, ctypes.byref(n), ctypes.byref(xb)))
        self.n_real, self.exit_base = n.value, xb.value
        return self.n_real, self.exit_base

    def roots(self):
        out = np.zeros(max(self.n_real, 1), dtype=np.int32)
        self.check(self.lib.bdr_slab_roots(self.h, out.ctypes.data, out.shape[0]))
        return torch.as_tensor(out[:self.n_real], device=self.device)

    def first_voxel(self, n_slots):
        out = torch.empty(max(n_slots, 1), dtype=torch.int32, device=self.device)
        self._sync()
        self.check(self.lib.bdr_slab_first_voxel(self.h, int(n_slots), out.data_ptr()))
        return out[:n_slots]

    def apply_rank(self, rank_lut):
        rank_lut = rank_lut.contiguous()
        self.check(self.lib.bdr_slab_apply_rank(self.h, rank_lut.data_ptr()))

    def first_voxel_labels(self, n_labels):
        out = torch.empty(max(n_labels, 1), dtype=torch.int32, device=self.device)
        self._sync()
        self.check(self.lib.bdr_slab_first_voxel_labels(self.h, int(n_labels), out.data_ptr()))
        return out[:n_labels]

    def relabel(self, lut):
        lut = lut.contiguous()
        self._sync()
        self.check(self.lib.bdr_slab_relabel(self.h, 0, lut.data_ptr()))

    def comm_id(self):
        buf = ctypes.create_string_buffer(128)
        self.check(self.lib.bdr_slab_comm_id(buf))
        return buf.raw

    def comm_init(self, world, rank, ident):
        self.check(self.lib.bdr_slab_comm_init(self.h, int(world), int(rank), ident))

    def exchange(self, what):
        self.check(self.lib.bdr_slab_exchange(self.h, int(what)))

    def rounds_native(self, dist_mat, T_grad, max_passes=64, cap=256):
        d = np.ascontiguousarray(dist_mat, dtype=np.float64)
        t = np.ascontiguousarray(T_grad, dtype=np.float64)
        hist = np.zeros((cap, 2), dtype=np.int64)
        n, settled = ctypes.c_int64(0), ctypes.c_int(0)
        self.check(self.lib.bdr_slab_rounds(self.h, 0, d.ctypes.data, t.ctypes.data, int(max_passes),
                                            hist.ctypes.data, cap, ctypes.byref(n), ctypes.byref(settled)))
        return [tuple(int(x) for x in r) for r in hist[:min(n.value, cap)]], bool(settled.value)

    def refine_native(self, mode, iters, dist_mat, T_grad, cap=256):
        d = np.ascontiguousarray(dist_mat, dtype=np.float64)
        t = np.ascontiguousarray(T_grad, dtype=np.float64)
        hist = np.zeros((cap, 2), dtype=np.int64)
        n = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_refine(self.h, 0, 0 if mode == 'all' else 1, int(iters), d.ctypes.data,
                                            t.ctypes.data, ctypes.byref(n), hist.ctypes.data, cap))
        return [tuple(int(x) for x in r) for r in hist[:min(n.value, cap)]]

    def known(self):
        if self._known is None:
            self._known = torch.as_tensor(_DevArray(self._ptr(5), self.shape, '|i1'), device=self.device)
        return self._known

    def ec_begin(self):
        self.check(self.lib.bdr_slab_ec_begin(self.h, 0))

    def ec_round(self):
        u = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_ec_round(self.h, ctypes.byref(u)))
        return u.value

    def ec_finish(self):
        e = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_ec_finish(self.h, 0, ctypes.byref(e)))
        return e.value

    def edge_pass(self):
        e = ctypes.c_int64(0)
        self.check(self.lib.bdr_edge_pass(self.h, 0, ctypes.byref(e)))
        return e.value

    def first_pass(self):
        self._sync()
        e = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_first_pass(self.h, 0, ctypes.byref(e)))
        return e.value

    def trace(self, dist_mat, T_grad, want_list):
        self._sync()
        d = np.ascontiguousarray(dist_mat, dtype=np.float64)
        t = np.ascontiguousarray(T_grad, dtype=np.float64)
        ch = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_trace(self.h, 0, d.ctypes.data, t.ctypes.data, int(want_list),
                                           ctypes.byref(ch)))
        return ch.value

    def requeue(self, extra):
        extra = extra.contiguous()
        self._sync()
        q = ctypes.c_int64(0)
        self.check(self.lib.bdr_slab_requeue(self.h, 0, extra.data_ptr() if extra.numel() else None,
                                             int(extra.numel()), ctypes.byref(q)))
        return q.value

    def trace_pass(self, dist_mat, T_grad, want_list=False):
        d = np.ascontiguousarray(dist_mat, dtype=np.float64)
        t = np.ascontiguousarray(T_grad, dtype=np.float64)
        ch, esc = ctypes.c_int64(0), ctypes.c_int64(0)
        self.check(self.lib.bdr_trace_pass_list(self.h, 0, d.ctypes.data, t.ctypes.data, int(want_list),
                                                ctypes.byref(ch), ctypes.byref(esc)))
        return ch.value, esc.value

    def charge_sum(self, n, dV, which_density=0):
        q, v = np.zeros(n), np.zeros(n)
        self._sync()
        self.check(self.lib.bdr_charge_sum(self.h, 0, which_density, float(dV), n, q.ctypes.data,
                                           v.ctypes.data))
        return q, v

    def timer_start(self):
        self.check(self.lib.bdr_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_double(0)
        self.check(self.lib.bdr_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        n = ctypes.c_int64(0)
        self.check(self.lib.bdr_launch_count(self.h, ctypes.byref(n)))
        return n.value


# ---------------------------------------------------------------------------
def bench(args, rank, world, local):
    """bench.py's N > 1 arm: weak scaling, 2^30 voxels per GPU, x-slabs."""
    import json
    import os
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B
    from . import build, geometry as geo, synth
    build.build()
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', local))
    comm = Comm()
    shape = B.SHAPES[world] if not args.size else (args.size * world, args.size, args.size)
    case, cells = B.workload_case(shape)
    dm = geo.distance_matrix(case['lattice'], shape)
    T = geo.T_grad(case['lattice'], shape)
    sb = ShardedBader(shape, comm, lambda ws, h: SlabBackend(ws, h, device=local), halo=args.halo)
    tx, ty, tz = synth.separable_tables(case)
    sb.backend.synth_separable(0, np.ascontiguousarray(tx[:, sb.window_x]), ty, tz)
    N = int(np.prod(shape))

    def step():
        sb.backend.clear_labels()
        sb.neargrid(dm, T)                 # ongrid seed + exits + numbering + rounds to quiescence
        return sb.refine(dm, T, 2, mode='changed')   # the caller's refine(), same mode as one GPU

    be = sb.backend
    dev = torch.device('cuda', local)
    for _ in range(args.warmup):
        hist = step()
    sb.phase_ms = {}
    clocks = B.ClockSampler(local)
    if rank == 0:
        clocks.start()
    be.check(be.lib.bdr_profile_enable(be.h, 1))
    be.check(be.lib.bdr_profile_reset(be.h))
    torch.cuda.synchronize()
    dist.barrier()
    l0 = be.launch_count()
    be.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hist = step()
    ms = be.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    wall = (time.perf_counter() - t0) * 1e3
    launches = be.launch_count() - l0
    prof = {}
    for i, name in enumerate(B_FAMILIES()):
        pm, pn = ctypes.c_double(0), ctypes.c_int64(0)
        be.check(be.lib.bdr_profile_get(be.h, i, ctypes.byref(pm), ctypes.byref(pn)))
        if pn.value:
            prof[name] = (pm.value, pn.value)
    ts, tv = ctypes.c_int64(0), ctypes.c_int64(0)
    be.check(be.lib.bdr_trace_steps(be.h, ctypes.byref(ts), ctypes.byref(tv)))
    be.check(be.lib.bdr_profile_enable(be.h, 0))
    clock_info = clocks.stop() if rank == 0 else None
    ms = comm.allreduce_max(ms, dev)
    launches = comm.allreduce_sum(launches, dev)
    ksum = sum(v[0] for v in prof.values()) / args.steps
    ksum_max = comm.allreduce_max(ksum, dev)           # slowest rank's kernels: the load imbalance
    ktrace_max = comm.allreduce_max(prof.get('trace', (0.0, 0))[0] / args.steps, dev)
    ms_per_step = ms / args.steps

    # ---- e2e: every rank uploads its slab window from pinned host memory, runs the
    # step and reads its labels back (narrowed like thread_handlers.bader_calc does)
    e2e = None
    if not args.no_e2e:
        from .utils import dtype_calc
        ldt = np.dtype(dtype_calc(-max(int(sb.maxima.shape[0]), 1)))
        nwin = be.N
        try:
            host_rho = torch.empty(nwin, dtype=torch.float64, pin_memory=True).numpy()
            host_lab = torch.empty(nwin * ldt.itemsize, dtype=torch.uint8, pin_memory=True).numpy()
            ok = 1
        except RuntimeError:
            ok = 0
        if comm.allreduce_sum(ok, dev) != world:
            args.no_e2e = True
    if not args.no_e2e:
        be.check(be.lib.bdr_download_density(be.h, 0, host_rho.ctypes.data))

        def e2e_step():
            be._sync()
            be.check(be.lib.bdr_upload_density(be.h, 0, host_rho.ctypes.data))
            h = step()
            be._sync()
            be.check(be.lib.bdr_download_labels(be.h, 0, host_lab.ctypes.data, ldt.itemsize))
            return h

        e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        dt = comm.allreduce_max(dt, dev)
        e2e = {"value": N / dt, "unit": B.UNIT, "h2d_bytes_per_step": nwin * 8 * world,
               "d2h_bytes_per_step": nwin * ldt.itemsize * world, "ms_per_step": dt * 1e3,
               "api": "per rank: bdr_upload_density(window, pinned host) + sharded step + "
                      "bdr_download_labels(narrowed); max over ranks"}
        del host_rho, host_lab
    if rank == 0 and getattr(sb, 'phase_ms', None):
        nst = args.steps * (1 if args.no_e2e else 2) + (0 if args.no_e2e else 1)
        print("[sharded] phases, ms per step (BDR_PHASES): " +
              ", ".join(f"{k} {v / nst:.2f}" for k, v in sb.phase_ms.items()), file=sys.stderr)
    if rank == 0:
        kernels, roofline = B.kernel_accounting(prof, be.N, args.steps, ts.value, tv.value, ms_per_step)
        roofline["note"] = "rank 0's kernels on its slab window; the step also holds NCCL exchanges"
        out = {
            "metric": B.METRIC, "value": N / (ms_per_step * 1e-3), "unit": B.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": B.workload_name(shape), "atoms": len(case['amps']),
                       "maxima": int(sb.maxima.shape[0]), "method": "neargrid",
                       "refine_method": "neargrid", "refine_mode": ["changed", 2],
                       "parallelism": f"{world} x-slabs, halo {args.halo} planes, NCCL ring exchange "
                                      f"of halo planes / exit labels, NVLink peer loads in the trace",
                       "l2": "per-GPU inputs are far larger than the 126 MB L2; no flush"},
            "roofline": roofline, "cpu_baseline": None,
            "e2e": e2e, "gpu_launches": launches, "clocks": clock_info,
            "wall_ms_per_step": wall / args.steps, "kernels": kernels,
            "refine_history_last_step": hist,
            "kernel_ms_per_step": {"rank0": ksum, "max_over_ranks": ksum_max, "trace_max_over_ranks": ktrace_max},
            "exit_rounds": sb.exit_rounds, "neargrid_passes": len(sb.neargrid_history),
            "neargrid_settled": sb.settled,
        }
        print(json.dumps(out))
    dist.destroy_process_group()


def B_FAMILIES():
    from .engine import FAMILIES
    return FAMILIES
