"""Drop-in replacements for the jitted helpers `pybader.interface` imports from
`pybader.utils` (interface.py:18): same names and positional signatures.

    dtype_calc      utils.py:15-37   (host logic, restated)
    vacuum_assign   utils.py:383-401
    charge_sum      utils.py:236-252
    atom_assign     utils.py:186-232
    volume_mask     utils.py:462-476
"""
import numpy as np

from . import session, sharded_handlers
from .engine import LABELS_BADER, RHO_CHARGE, Engine


def dtype_calc(max_val):
    """Smallest integer dtype name able to hold max_val (negative -> signed,
    with the reference's factor 2 headroom)."""
    signed = max_val < 0
    if signed:
        max_val *= -2
    names = ['int8', 'int16', 'int32', 'int64'] if signed else ['uint8', 'uint16', 'uint32', 'uint64']
    if max_val <= 255:
        return names[0]
    if max_val <= 65535:
        return names[1]
    if max_val <= 4294967295:
        return names[2]
    return names[3]


def vacuum_assign(reference, volumes, vac_tol, density, voxel_volume):
    if sharded_handlers.active():
        return sharded_handlers.vacuum_assign(reference, volumes, vac_tol, density, voxel_volume)
    s = session.get(reference.shape)
    s.reference(reference)
    dslot = s.density_slot(density, prefer=RHO_CHARGE)
    # fresh (all-zero) labels are cleared on the device, labelled ones (bader-read's
    # re-threshold flow, entry_points.py:238-255) uploaded; the content hash that keys
    # residency tells which, so the host array is read once
    s.label_slot(volumes, force=LABELS_BADER)
    charge, volume, count = s.engine.vacuum_assign(vac_tol, voxel_volume, dslot, want_count=True)
    if count:
        s.labels_to_host(LABELS_BADER, out=volumes)   # only then did any label change
    return volumes, charge, volume


def charge_sum(charge, volume, voxel_volume, density, volumes):
    if sharded_handlers.active():
        return sharded_handlers.charge_sum(charge, volume, voxel_volume, density, volumes)
    s = session.get(volumes.shape)
    lslot = s.label_slot(volumes, prefer=LABELS_BADER)
    # first free slot: reference, then charge, then spin (so charge and spin
    # both stay resident next to the reference)
    dslot = s.density_slot(density, prefer=s.free_density_slot())
    s.engine.charge_sum(lslot, dslot, voxel_volume, charge, volume)


def atom_assign(bader_max, atoms, lattice, i_c=None):
    e = Engine((1, 1, 1), session._device)
    try:
        e.clear_labels(LABELS_BADER)
        return e.assign_atoms(bader_max, atoms, lattice)
    finally:
        e.close()


def volume_mask(volumes, density, vol_num):
    if sharded_handlers.active():
        return sharded_handlers.volume_mask(volumes, density, vol_num)
    s = session.get(volumes.shape)
    lslot = s.label_slot(volumes, prefer=LABELS_BADER)
    dslot = s.density_slot(density, prefer=s.free_density_slot())
    return s.engine.volume_mask(lslot, dslot, vol_num)
