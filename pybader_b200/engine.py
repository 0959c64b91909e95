"""`Engine`: one device handle of libbader_b200.so behind numpy arguments.

Thin, stateful mirror of the C ABI; the reference-shaped functions live in
`pybader_b200.thread_handlers` / `pybader_b200.utils`.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check

RHO_REFERENCE, RHO_CHARGE, RHO_SPIN = 0, 1, 2
LABELS_BADER, LABELS_ATOMS = 0, 1
METHODS = {'ongrid': 0, 'neargrid': 1}          # methods.__contains__, methods.py:12
REFINE_METHODS = {'neargrid': 1}                # refinement.__contains__, refinement.py:13
MODES = {'all': 0, 'changed': 1}

FAMILIES = ['vacuum', 'stencil', 'resolve', 'relabel', 'edge_flag', 'edge_dilate', 'trace',
            'edge_check', 'charge_sum', 'assign', 'surface', 'narrow', 'synth', 'first',
            'edge_confirm', 'trace_peer', 'edge_eq']


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Engine:
    def __init__(self, shape, device=0):
        self.lib = _lib.load()
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) != 3:
            raise ValueError("density must be rank 3")
        self.N = int(np.prod(self.shape))
        h = ctypes.c_void_p()
        check(self.lib.bdr_create(int(device), *self.shape, ctypes.byref(h)))
        self.h = h
        self.device = device
        self.n_max = 0

    def close(self):
        if getattr(self, 'h', None):
            self.lib.bdr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement -------------------------------------------------------
    def upload_density(self, which, rho):
        rho = _f64(rho)
        if rho.shape != self.shape:
            raise ValueError(f"density shape {rho.shape} != engine shape {self.shape}")
        check(self.lib.bdr_upload_density(self.h, which, _ptr(rho)))

    def download_density(self, which):
        out = np.empty(self.shape, dtype=np.float64)
        check(self.lib.bdr_download_density(self.h, which, _ptr(out)))
        return out

    def alias_density(self, which, of):
        check(self.lib.bdr_alias_density(self.h, which, of))

    def copy_density(self, dst, src):
        check(self.lib.bdr_copy_density(self.h, dst, src))

    def upload_labels(self, which, labels):
        labels = np.ascontiguousarray(labels)
        if labels.shape != self.shape:
            raise ValueError("label shape mismatch")
        if labels.dtype.kind not in 'iu' or labels.dtype.itemsize not in (1, 2, 4, 8):
            raise TypeError(f"unsupported label dtype {labels.dtype}")
        if labels.dtype.kind == 'u':
            labels = labels.astype(np.int64)
        check(self.lib.bdr_upload_labels(self.h, which, _ptr(labels), labels.dtype.itemsize))

    def download_labels(self, which, dtype=np.int32, out=None):
        dtype = np.dtype(dtype)
        if out is None:
            out = np.empty(self.shape, dtype=dtype)
        assert out.flags.c_contiguous and out.dtype == dtype and out.shape == self.shape
        check(self.lib.bdr_download_labels(self.h, which, _ptr(out), dtype.itemsize))
        return out

    def download_known(self):
        out = np.empty(self.shape, dtype=np.int8)
        check(self.lib.bdr_download_known(self.h, _ptr(out)))
        return out

    def clear_labels(self, which=LABELS_BADER):
        check(self.lib.bdr_clear_labels(self.h, which))

    # -- hot path ------------------------------------------------------------
    def vacuum_assign(self, vac_tol, voxel_volume, which_density=RHO_REFERENCE, want_count=False):
        q, v = ctypes.c_double(0), ctypes.c_double(0)
        check(self.lib.bdr_vacuum_assign(self.h, float(vac_tol), float(voxel_volume),
                                         which_density, ctypes.byref(q), ctypes.byref(v)))
        if not want_count:
            return q.value, v.value
        n = ctypes.c_int64(0)
        check(self.lib.bdr_vacuum_count(self.h, ctypes.byref(n)))
        return q.value, v.value, n.value

    def bader_calc(self, method, dist_mat, T_grad):
        d, t = _f64(dist_mat), _f64(T_grad)
        n = ctypes.c_int64(0)
        check(self.lib.bdr_bader_calc(self.h, METHODS[method], _ptr(d), _ptr(t), ctypes.byref(n)))
        self.n_max = n.value
        return self.maxima()

    def maxima(self):
        out = np.zeros((self.n_max, 3), dtype=np.int64)
        check(self.lib.bdr_get_maxima(self.h, _ptr(out), self.n_max))
        return out

    def refine(self, which, mode, iters, dist_mat, T_grad, hist_cap=256):
        d, t = _f64(dist_mat), _f64(T_grad)
        run = ctypes.c_int64(0)
        hist = np.zeros((hist_cap, 2), dtype=np.int64)
        check(self.lib.bdr_refine(self.h, which, MODES[mode.lower()], int(iters), _ptr(d), _ptr(t),
                                  ctypes.byref(run), _ptr(hist), hist_cap))
        return [tuple(int(x) for x in row) for row in hist[:min(run.value, hist_cap)]]

    def edge_find(self, which=LABELS_BADER):
        e = ctypes.c_int64(0)
        check(self.lib.bdr_edge_find(self.h, which, ctypes.byref(e)))
        return e.value

    def charge_sum(self, which_labels, which_density, voxel_volume, charge, volume):
        assert charge.dtype == np.float64 and volume.dtype == np.float64
        assert charge.flags.c_contiguous and volume.flags.c_contiguous
        check(self.lib.bdr_charge_sum(self.h, which_labels, which_density, float(voxel_volume),
                                      charge.shape[0], _ptr(charge), _ptr(volume)))

    def assign_atoms(self, maxima_cart, atoms_cart, lattice):
        m, a, l = _f64(maxima_cart).reshape(-1, 3), _f64(atoms_cart).reshape(-1, 3), _f64(lattice)
        who = np.zeros(m.shape[0], dtype=np.int64)
        dist = np.zeros(m.shape[0], dtype=np.float64)
        check(self.lib.bdr_assign_atoms(self.h, _ptr(m), m.shape[0], _ptr(a), a.shape[0], _ptr(l),
                                        _ptr(who), _ptr(dist)))
        return who, dist

    def surface_distance(self, which, lattice, atoms_cart):
        a, l = _f64(atoms_cart).reshape(-1, 3), _f64(lattice)
        dist = np.zeros(a.shape[0], dtype=np.float64)
        found = ctypes.c_int(0)
        check(self.lib.bdr_surface_distance(self.h, which, _ptr(l), _ptr(a), a.shape[0],
                                            _ptr(dist), ctypes.byref(found)))
        return dist if found.value else None

    def volume_mask(self, which_labels, which_density, vol_num):
        out = np.empty(self.shape, dtype=np.float64)
        check(self.lib.bdr_volume_mask(self.h, which_labels, which_density, int(vol_num), _ptr(out)))
        return out

    def run(self, rho, vac_tol, voxel_volume, method, refine_mode, refine_iters, dist_mat, T_grad,
            label_dtype=np.int32, max_cap=1 << 16, want_sums=True, out_labels=None):
        """One-shot host-in / host-out pipeline (bdr_run)."""
        rho = _f64(rho)
        d, t = _f64(dist_mat), _f64(T_grad)
        label_dtype = np.dtype(label_dtype)
        labels = out_labels if out_labels is not None else np.empty(self.shape, dtype=label_dtype)
        maxima = np.zeros((max_cap, 3), dtype=np.int64)
        charge = np.zeros(max_cap, dtype=np.float64)
        volume = np.zeros(max_cap, dtype=np.float64)
        n = ctypes.c_int64(0)
        tol = float('nan') if vac_tol is None else float(vac_tol)
        check(self.lib.bdr_run(self.h, _ptr(rho), tol, float(voxel_volume), METHODS[method],
                               MODES[refine_mode.lower()], int(refine_iters), _ptr(d), _ptr(t),
                               _ptr(labels), label_dtype.itemsize, ctypes.byref(n), _ptr(maxima),
                               max_cap, _ptr(charge) if want_sums else None,
                               _ptr(volume) if want_sums else None))
        self.n_max = n.value
        return labels, maxima[:n.value].copy(), charge[:n.value].copy(), volume[:n.value].copy()

    # -- measurement -----------------------------------------------------------
    def profile(self, on=True):
        check(self.lib.bdr_profile_enable(self.h, int(on)))

    def profile_reset(self):
        check(self.lib.bdr_profile_reset(self.h))

    def profile_get(self):
        out = {}
        for i, name in enumerate(FAMILIES):
            ms, n = ctypes.c_double(0), ctypes.c_int64(0)
            check(self.lib.bdr_profile_get(self.h, i, ctypes.byref(ms), ctypes.byref(n)))
            if n.value:
                out[name] = (ms.value, n.value)
        return out

    def launch_count(self):
        n = ctypes.c_int64(0)
        check(self.lib.bdr_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def sync_count(self):
        n = ctypes.c_int64(0)
        check(self.lib.bdr_sync_count(self.h, ctypes.byref(n)))
        return n.value

    def timer_start(self):
        check(self.lib.bdr_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_double(0)
        check(self.lib.bdr_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def trace_steps(self):
        s, v = ctypes.c_int64(0), ctypes.c_int64(0)
        check(self.lib.bdr_trace_steps(self.h, ctypes.byref(s), ctypes.byref(v)))
        return s.value, v.value

    def synchronize(self):
        check(self.lib.bdr_synchronize(self.h))

    # -- synthetic inputs ------------------------------------------------------
    def synth_separable(self, which, tx, ty, tz):
        tx, ty, tz = _f64(tx), _f64(ty), _f64(tz)
        assert tx.shape[1] == self.shape[0] and ty.shape[1] == self.shape[1] and tz.shape[1] == self.shape[2]
        check(self.lib.bdr_synth_separable(self.h, which, _ptr(tx), _ptr(ty), _ptr(tz), tx.shape[0]))

    def synth_general(self, which, lattice, frac_atoms, amps, sigmas):
        l, f, a, s = _f64(lattice), _f64(frac_atoms), _f64(amps), _f64(sigmas)
        check(self.lib.bdr_synth_general(self.h, which, _ptr(l), _ptr(f), _ptr(a), _ptr(s), f.shape[0]))

    def set_option(self, option, value):
        check(self.lib.bdr_set_option(self.h, int(option), int(value)))

    def device_ptr(self, what):
        p = ctypes.c_void_p()
        check(self.lib.bdr_device_ptr(self.h, what, ctypes.byref(p)))
        return p.value
