"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Restates, for ``threads=1``, the host-side logic of the reference's dispatch
layer (pybader/thread_handlers.py) on top of ``bader_oracle.c``:

* ``bader_calc``        thread_handlers.py:15-75  (one brick; labels made
                        0-based by utils.volume_offset, utils.py:497-510, then
                        narrowed with utils.dtype_calc, utils.py:15-37)
* ``refine``            thread_handlers.py:128-236
* ``assign_to_atoms``   thread_handlers.py:78-125
* ``surface_distance``  thread_handlers.py:239-297
* ``vacuum_assign`` / ``charge_sum``  utils.py:383-401 / 236-252

The function names and positional signatures are the reference's so that the
parity tests read like calls into ``pybader.thread_handlers``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbader_oracle.so")
_lib = None

_i64 = ctypes.c_int64
_f64 = ctypes.c_double
_p = ctypes.c_void_p


def build(force=False):
    """Compile bader_oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "bader_oracle.c")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_vacuum_assign.argtypes = [_p, _p, _i64, _f64, _p, _f64, _p, _p]
        L.orc_vacuum_assign.restype = None
        L.orc_ongrid.argtypes = [_p, _p, _i64, _i64, _i64, _p, _p, _i64]
        L.orc_ongrid.restype = _i64
        L.orc_neargrid.argtypes = [_p, _p, _i64, _i64, _i64, _p, _p, _p, _i64]
        L.orc_neargrid.restype = _i64
        L.orc_edge_find.argtypes = [_p, _p, _p, _i64, _i64, _i64]
        L.orc_edge_find.restype = _i64
        L.orc_edge_check.argtypes = [_p, _p, _p, _i64, _i64, _i64, _p, _p]
        L.orc_edge_check.restype = None
        L.orc_refine_neargrid.argtypes = [_p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _i64]
        L.orc_refine_neargrid.restype = _i64
        L.orc_charge_sum.argtypes = [_p, _p, _i64, _f64, _p, _p, _i64]
        L.orc_charge_sum.restype = None
        L.orc_atom_assign.argtypes = [_p, _i64, _p, _i64, _p, _p, _p]
        L.orc_atom_assign.restype = None
        L.orc_volume_assign.argtypes = [_p, _i64, _p]
        L.orc_volume_assign.restype = None
        L.orc_surface_dist.argtypes = [_p, _p, _i64, _i64, _i64, _p, _p, _i64, _p]
        L.orc_surface_dist.restype = None
        L.orc_volume_mask.argtypes = [_p, _p, _i64, ctypes.c_int32, _p]
        L.orc_volume_mask.restype = None
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def dtype_calc(max_val):
    """utils.dtype_calc (utils.py:15-37): smallest int dtype that holds max_val."""
    names = (['int8', 'int16', 'int32', 'int64'] if max_val < 0
             else ['uint8', 'uint16', 'uint32', 'uint64'])
    if max_val < 0:
        max_val *= -2
    for lim, name in zip((255, 65535, 4294967295), names):
        if max_val <= lim:
            return name
    return names[3]


# --------------------------------------------------------------------------
def vacuum_assign(reference, volumes, vac_tol, density, voxel_volume):
    ref, den = _f(reference), _f(density)
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    q, v = _f64(0), _f64(0)
    lib().orc_vacuum_assign(_ptr(ref), _ptr(v32), v32.size, float(vac_tol), _ptr(den),
                            float(voxel_volume), ctypes.byref(q), ctypes.byref(v))
    volumes[...] = v32
    return volumes, q.value, v.value


def raw_method(method, density, volumes, dist_mat, T_grad):
    """methods.ongrid / methods.neargrid on the whole volume.  Returns
    (volumes int32 with 1-based labels, maxima int64[n,3])."""
    rho = _f(density)
    v32 = np.ascontiguousarray(volumes, dtype=np.int32).copy()
    dist, T = _f(dist_mat), _f(T_grad)
    cap = 1 << 12
    while True:
        work = v32.copy()
        maxima = np.zeros((cap, 3), dtype=np.int64)
        if method == 'ongrid':
            n = lib().orc_ongrid(_ptr(rho), _ptr(work), *rho.shape, _ptr(dist),
                                 _ptr(maxima), cap)
        elif method == 'neargrid':
            n = lib().orc_neargrid(_ptr(rho), _ptr(work), *rho.shape, _ptr(dist),
                                   _ptr(T), _ptr(maxima), cap)
        else:
            raise AttributeError(method)     # getattr(methods, method) would
        if n >= 0:
            return work, maxima[:n].copy()
        cap *= 16


def bader_calc(method, density, volumes, dist_mat, T_grad, threads=1):
    work, maxima = raw_method(method, density, volumes, dist_mat, T_grad)
    work[work > 0] -= 1                      # utils.volume_offset, one brick
    out = work.astype(dtype_calc(-maxima.shape[0]))
    return maxima, out


def edge_find(known, density, volumes):
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    rho = _f(density)
    assert known.dtype == np.int8 and known.flags.c_contiguous
    return lib().orc_edge_find(_ptr(known), _ptr(rho), _ptr(v32), *rho.shape)


def edge_check(known, density, volumes):
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    rho = _f(density)
    assert known.dtype == np.int8 and known.flags.c_contiguous
    c, e = _i64(0), _i64(0)
    lib().orc_edge_check(_ptr(known), _ptr(rho), _ptr(v32), *rho.shape,
                         ctypes.byref(c), ctypes.byref(e))
    return c.value, e.value


def refine_neargrid(known, rknown, density, volumes, dist_mat, T_grad,
                    step_cap=1 << 20):
    """refinement.neargrid on the whole volume; volumes modified in place."""
    rho, dist, T = _f(density), _f(dist_mat), _f(T_grad)
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    own = v32 is volumes
    ch = lib().orc_refine_neargrid(_ptr(known), _ptr(rknown), _ptr(rho), _ptr(v32),
                                   *rho.shape, _ptr(dist), _ptr(T), step_cap)
    if ch < 0:
        raise RuntimeError("oracle: trajectory exceeded step cap")
    if not own:
        volumes[...] = v32
    return ch


def refine(method, refine_mode, density, volumes, dist_mat, T_grad, threads=1,
           log=None):
    """thread_handlers.refine; `log` (a list) receives (edges, changed) per
    iteration."""
    if method != 'neargrid':                 # getattr(refinement, method) fails
        return
    check_mode, iters = tuple(refine_mode)
    if iters == 0:
        return
    known = np.zeros(density.shape, dtype=np.int8)
    edges = edge_find(known, density, volumes)
    if edges == 0:
        return
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    if v32 is volumes:
        v32 = volumes
    rknown = known.copy()
    changed = refine_neargrid(known, rknown, density, v32, dist_mat, T_grad)
    if log is not None:
        log.append((edges, changed))
    if iters < 0:
        iters = float('inf')
    it = 2
    while it <= iters:
        if check_mode.lower() == 'all':
            known = np.zeros(density.shape, dtype=np.int8)
            edges = edge_find(known, density, v32)
        else:
            _, edges = edge_check(known, density, v32)
        rknown = known.copy()
        changed = refine_neargrid(known, rknown, density, v32, dist_mat, T_grad)
        if log is not None:
            log.append((edges, changed))
        if changed == 0:
            break
        it += 1
    if v32 is not volumes:
        volumes[...] = v32


def charge_sum(charge, volume, voxel_volume, density, volumes):
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    den = _f(density)
    assert charge.dtype == np.float64 and volume.dtype == np.float64
    lib().orc_charge_sum(_ptr(charge), _ptr(volume), charge.shape[0], float(voxel_volume),
                         _ptr(den), _ptr(v32), v32.size)


def atom_assign(bader_max, atoms, lattice):
    b, a, l = _f(bader_max), _f(atoms), _f(lattice)
    who = np.zeros(b.shape[0], dtype=np.int64)
    dist = np.zeros(b.shape[0], dtype=np.float64)
    lib().orc_atom_assign(_ptr(b), b.shape[0], _ptr(a), a.shape[0], _ptr(l),
                          _ptr(who), _ptr(dist))
    return who, dist


def assign_to_atoms(bader_max, atoms, lattice, volumes, threads=1):
    who, dist = atom_assign(bader_max, atoms, lattice)
    v32 = np.ascontiguousarray(volumes, dtype=np.int32).copy()
    lib().orc_volume_assign(_ptr(v32), v32.size, _ptr(who))
    return who, dist, v32.astype(dtype_calc(-np.asarray(atoms).shape[0]))


def surface_distance(density, volumes, lattice, atoms, threads=1):
    known = np.zeros(volumes.shape, dtype=np.int8)
    edges = edge_find(known, density, volumes)
    if edges == 0:
        return None
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    l, a = _f(lattice), _f(atoms)
    dist = np.zeros(a.shape[0], dtype=np.float64)
    lib().orc_surface_dist(_ptr(known), _ptr(v32), *v32.shape, _ptr(l), _ptr(a),
                           a.shape[0], _ptr(dist))
    # thread_handlers.py:289-297 with a single brick: zeros mean "no edge"
    out = np.zeros(a.shape[0], dtype=np.float64)
    d = dist.copy()
    d[d == 0] = np.inf
    sel = np.isfinite(d)
    out[sel] = d[sel]
    return out


def volume_mask(volumes, density, vol_num):
    v32 = np.ascontiguousarray(volumes, dtype=np.int32)
    den = _f(density)
    out = np.zeros(den.shape, dtype=np.float64)
    lib().orc_volume_mask(_ptr(v32), _ptr(den), v32.size, int(vol_num), _ptr(out))
    return out
