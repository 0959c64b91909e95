"""Drop-in replacements for `pybader.thread_handlers` (same names, same
positional signatures, same return values) that run on the B200 through the
C ABI instead of the numba jits.

    bader_calc        thread_handlers.py:15-75
    assign_to_atoms   thread_handlers.py:78-125
    refine            thread_handlers.py:128-236
    surface_distance  thread_handlers.py:239-297

`threads` is accepted and ignored (there is no CPU threading here).  Under `torchrun`
(WORLD_SIZE > 1) every entry point runs sharded over the ranks' GPUs
(pybader_b200.sharded_handlers) and returns the same arrays on every rank.
"""
import numpy as np

from . import session, sharded_handlers
from .engine import LABELS_ATOMS, LABELS_BADER, METHODS, MODES, REFINE_METHODS
from .utils import dtype_calc


def bader_calc(method, density, volumes, dist_mat, T_grad, threads=1):
    if sharded_handlers.active():       # WORLD_SIZE > 1: x-slabs over the ranks' GPUs
        return sharded_handlers.bader_calc(method, density, volumes, dist_mat, T_grad, threads)
    if method not in METHODS:
        # the reference does getattr(methods, method) (thread_handlers.py:26)
        raise AttributeError(f"module 'pybader.methods' has no attribute '{method}'")
    s = session.get(density.shape)
    s.reference(density)
    s.label_slot(volumes, force=LABELS_BADER)
    bader_max = s.engine.bader_calc(method, dist_mat, T_grad)
    out = s.labels_to_host(LABELS_BADER, dtype_calc(-bader_max.shape[0]))
    return bader_max, out


def refine(method, refine_mode, density, volumes, dist_mat, T_grad, threads=1):
    if sharded_handlers.active():
        sharded_handlers.refine(method, refine_mode, density, volumes, dist_mat, T_grad, threads)
        refine.last_history = sharded_handlers.refine.last_history
        return
    if method not in REFINE_METHODS:
        return  # getattr(refinement, method) fails -> silently no refinement (l.140-143)
    check_mode, iters = tuple(refine_mode)
    if iters == 0:
        return
    mode = 'all' if check_mode.lower() == 'all' else 'changed'
    s = session.get(density.shape)
    s.reference(density)
    slot = s.label_slot(volumes, prefer=LABELS_BADER)
    history = s.engine.refine(slot, mode, iters, dist_mat, T_grad)
    if history and any(ch for _, ch in history):
        s.labels_to_host(slot, out=volumes)
    refine.last_history = history


refine.last_history = []


def assign_to_atoms(bader_max, atoms, lattice, volumes, threads=1):
    if sharded_handlers.active():
        return sharded_handlers.assign_to_atoms(bader_max, atoms, lattice, volumes, threads)
    s = session.get(volumes.shape)
    s.label_slot(volumes, force=LABELS_BADER)
    bader_atoms, bader_distance = s.engine.assign_atoms(bader_max, atoms, lattice)
    atoms_volumes = s.labels_to_host(LABELS_ATOMS, dtype_calc(-np.asarray(atoms).shape[0]))
    return bader_atoms, bader_distance, atoms_volumes


def surface_distance(density, volumes, lattice, atoms, threads=1):
    if sharded_handlers.active():
        return sharded_handlers.surface_distance(density, volumes, lattice, atoms, threads)
    s = session.get(density.shape)
    s.reference(density)
    slot = s.label_slot(volumes, prefer=LABELS_ATOMS)
    return s.engine.surface_distance(slot, lattice, atoms)
