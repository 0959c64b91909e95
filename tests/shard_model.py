"""CPU model of the per-rank kernels (test infrastructure): numpy pointer field
+ pointer jumping for the seed, the C oracle for the refinement passes.  Lets
tests drive pybader_b200.sharded.ShardedBader over gloo without a GPU."""
import itertools

import numpy as np
import torch

from oracle import pyoracle as orc


def ongrid_pointers(rho, dist_mat):
    """steepest-ascent target (linear index) of every voxel of a periodic grid,
    same arithmetic and tie-break order as methods.py:87-117"""
    n = rho.size
    best = rho.copy()
    target = np.arange(n, dtype=np.int64).reshape(rho.shape)
    lin = np.arange(n, dtype=np.int64).reshape(rho.shape)
    for ix, iy, iz in itertools.product((-1, 0, 1), repeat=3):
        if ix == iy == iz == 0:
            continue
        w = dist_mat[ix, iy, iz]
        rn = np.roll(rho, (-ix, -iy, -iz), axis=(0, 1, 2))
        ln = np.roll(lin, (-ix, -iy, -iz), axis=(0, 1, 2))
        val = (rn - rho) * w
        val = val + rho
        upd = val > best
        best[upd] = val[upd]
        target[upd] = ln[upd]
    return target


class ModelBackend:
    def __init__(self, window_shape, halo, rho_window, vac_tol=None):
        self.shape = tuple(window_shape)
        self.halo = halo
        self.rho = np.ascontiguousarray(rho_window, dtype=np.float64)
        assert self.rho.shape == self.shape
        self.N = self.rho.size
        self.plane = self.shape[1] * self.shape[2]
        self.lab = np.zeros(self.shape, dtype=np.int32)
        if vac_tol is not None:
            self.lab[self.rho <= vac_tol] = -1
        self._labels_t = torch.from_numpy(self.lab)
        self.own_lo, self.own_hi = halo * self.plane, (self.shape[0] - halo) * self.plane
        self._known = np.zeros(self.shape, dtype=np.int8)
        self._known_t = torch.from_numpy(self._known)
        self.changed = np.zeros(0, dtype=np.int64)
        # global scan order of edge_check's centre selection: set by the test (ShardedBader knows it)
        self.x0w, self.NX = 0, self.shape[0]

    def known(self):
        return self._known_t

    def labels(self):
        return self._labels_t

    def seed(self, dist_mat, method='ongrid'):
        W = self.shape[0]
        ptr = ongrid_pointers(self.rho, np.asarray(dist_mat)).reshape(-1)
        lin = np.arange(self.N, dtype=np.int64)
        vac = self.lab.reshape(-1) == -1
        code = np.where(ptr == lin, np.int64(-10), ptr)          # roots marked, slots below
        code[vac] = -1
        xs = lin // self.plane
        exit_base = 2 * self.plane
        ex0, ex1 = xs == 0, xs == W - 1
        code[ex0] = -2 - (lin[ex0] % self.plane)
        code[ex1] = -2 - (self.plane + lin[ex1] % self.plane)
        roots = np.flatnonzero(code == -10)
        code[roots] = -2 - (exit_base + np.arange(roots.size))
        self._roots = roots.astype(np.int32)
        # pointer jumping
        for _ in range(64):
            p = code >= 0
            if not p.any():
                break
            nxt = code[code[p]]
            code[p] = nxt
        assert not (code >= 0).any()
        self.lab.reshape(-1)[:] = code.astype(np.int32)
        self.n_real = roots.size
        return self.n_real, exit_base

    def roots(self):
        return torch.from_numpy(self._roots.copy())

    def first_voxel(self, n_slots):
        out = np.full(n_slots, 0x7f7f7f7f, dtype=np.int32)
        flat = self.lab.reshape(-1)
        idx = np.arange(self.own_lo, self.own_hi)
        c = flat[self.own_lo:self.own_hi]
        sel = c <= -2
        s = -2 - c[sel].astype(np.int64)
        np.minimum.at(out, s, idx[sel].astype(np.int32))
        return torch.from_numpy(out)

    def apply_rank(self, rank_lut):
        lut = rank_lut.numpy()
        flat = self.lab.reshape(-1)
        sel = flat <= -2
        flat[sel] = lut[-2 - flat[sel].astype(np.int64)]

    def edge_pass(self):
        self._known[...] = 0
        orc.edge_find(self._known, self.rho, self.lab)
        return int((self._known.reshape(-1)[self.own_lo:self.own_hi] == -2).sum())

    def trace_pass(self, dist_mat, T_grad, want_list=False):
        before = self.lab.reshape(-1)[self.own_lo:self.own_hi].copy()
        orc.refine_neargrid(self._known, self._known.copy(), self.rho, self.lab, dist_mat, T_grad)
        after = self.lab.reshape(-1)[self.own_lo:self.own_hi]
        self.changed = np.flatnonzero(before != after) + self.own_lo if want_list else np.zeros(0, np.int64)
        return int((before != after).sum()), 0

    # ---- refinement.edge_check (refinement.py:409-508) in the phases the ranks exchange between;
    # a direct numpy restatement of the kernels k_ec_* (loops over the short changed list)
    def _unlin(self, v):
        nx, ny, nz = self.shape
        return v // (ny * nz), (v // nz) % ny, v % nz

    def _nbrs(self, v):
        nx, ny, nz = self.shape
        x, y, z = self._unlin(v)
        for ix, iy, iz in itertools.product((-1, 0, 1), repeat=3):
            yield ((x + ix) % nx, (y + iy) % ny, (z + iz) % nz)

    def _gidx(self, p):
        gx = (p[0] + self.x0w) % self.NX
        return (gx * self.shape[1] + p[1]) * self.shape[2] + p[2]

    def _classify(self, p):
        """0 not an edge, 1 edge and not a maximum, 2 edge and maximum (vacuum neighbours ignored)"""
        mine, here = self.lab[p], self.rho[p]
        e, m = False, True
        for q in self._nbrs((p[0] * self.shape[1] + p[1]) * self.shape[2] + p[2]):
            l = self.lab[q]
            if l == -1:
                continue
            e |= l != mine
            m &= not self.rho[q] > here
        return (2 if m else 1) if e else 0

    def ec_begin(self):
        for v in self.changed:
            p = self._unlin(int(v))
            if self._classify(p) == 2:
                self._known[p] = -4

    def ec_round(self):
        new, undecided = {}, 0
        for v in self.changed:
            p = self._unlin(int(v))
            if self._known[p] != -2:
                continue
            gv = self._gidx(p)
            out = blocked = False
            for q in self._nbrs(int(v)):
                if self._gidx(q) >= gv:
                    continue
                k = self._known[q]
                out |= k == -4
                blocked |= k == -2
            if out:
                new[p] = -5
            elif not blocked:
                new[p] = -4
            else:
                undecided += 1
        for p, k in new.items():
            self._known[p] = k
        return undecided

    def ec_finish(self):
        W, H = self.shape[0], self.halo
        centres = [self._unlin(int(v)) for v in self.changed if self._known[self._unlin(int(v))] == -4]
        for xs in (range(1, H), range(W - H, W - 1)):
            for x in xs:
                for y, z in zip(*np.nonzero(self._known[x] == -4)):
                    centres.append((x, int(y), int(z)))
        new_edges = []
        for c in centres:
            for pe in self._nbrs((c[0] * self.shape[1] + c[1]) * self.shape[2] + c[2]):
                cls = self._classify(pe)
                if cls == 0:
                    self._known[pe] = -1
                elif cls == 1 and self._known[pe] != -3:
                    self._known[pe] = -3
                    new_edges.append(pe)
        for e in new_edges:
            for q in self._nbrs((e[0] * self.shape[1] + e[1]) * self.shape[2] + e[2]):
                if self._known[q] >= 0:
                    self._known[q] = -1
        owned = 0
        for e in new_edges:
            self._known[e] = -2
            lin = (e[0] * self.shape[1] + e[1]) * self.shape[2] + e[2]
            owned += self.own_lo <= lin < self.own_hi
        for c in centres:
            if self._known[c] == -4:
                self._known[c] = -2
        return owned

    def charge_sum(self, n, dV, which_density=0):
        q, v = np.zeros(n), np.zeros(n)
        flat = self.lab.reshape(-1)[self.own_lo:self.own_hi]
        rho = self.rho.reshape(-1)[self.own_lo:self.own_hi]
        sel = flat >= 0
        np.add.at(q, flat[sel], rho[sel])
        np.add.at(v, flat[sel], 1.0)
        return q * dV, v * dV
