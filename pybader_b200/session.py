"""Device residency for the reference-shaped entry points.

The reference passes numpy arrays between stages (interface.py:449-534).  To
keep those signatures while not re-uploading a density for every stage, one
`Session` per grid shape remembers which host arrays its device buffers mirror.
An array is recognised by its address, shape, dtype and a strided sample of its
bytes; a mismatch simply re-uploads.
"""
import zlib

import numpy as np

from .engine import (LABELS_ATOMS, LABELS_BADER, RHO_CHARGE, RHO_REFERENCE, RHO_SPIN, Engine)

_sessions = {}
_device = 0


def set_device(device):
    """Select the CUDA device new sessions are created on."""
    global _device
    _device = int(device)


def fingerprint(a):
    a = np.asarray(a)
    flat = a.reshape(-1)
    step = max(1, flat.size // 4096)
    sample = np.ascontiguousarray(flat[::step])
    return (a.__array_interface__['data'][0], a.shape, a.dtype.str, a.strides,
            zlib.crc32(sample.tobytes()))


class Session:
    def __init__(self, shape):
        self.engine = Engine(shape, _device)
        self.shape = tuple(shape)
        self.rho_key = [None, None, None]
        self.label_key = [None, None]

    # densities -------------------------------------------------------------
    def density_slot(self, arr, prefer=RHO_CHARGE):
        """Slot holding `arr`, uploading into `prefer` if it is not resident."""
        key = fingerprint(arr)
        for slot in (RHO_REFERENCE, RHO_CHARGE, RHO_SPIN):
            if self.rho_key[slot] == key:
                return slot
        self.engine.upload_density(prefer, arr)
        self.rho_key[prefer] = key
        if prefer == RHO_REFERENCE:
            # stale aliases must not survive a new reference
            for slot in (RHO_CHARGE, RHO_SPIN):
                if self.rho_key[slot] is None:
                    self.engine.alias_density(slot, RHO_REFERENCE)
        return prefer

    def reference(self, arr):
        return self.density_slot(arr, prefer=RHO_REFERENCE)

    # labels ----------------------------------------------------------------
    def label_slot(self, arr, prefer=LABELS_BADER):
        key = fingerprint(arr)
        for slot in (LABELS_BADER, LABELS_ATOMS):
            if self.label_key[slot] == key:
                return slot
        self.engine.upload_labels(prefer, arr)
        self.label_key[prefer] = key
        return prefer

    def labels_to_host(self, slot, dtype=None, out=None):
        if out is not None:
            if out.flags.c_contiguous and out.dtype.kind == 'i':
                self.engine.download_labels(slot, out.dtype, out=out)
            else:
                out[...] = self.engine.download_labels(slot, np.int32)
            host = out
        else:
            host = self.engine.download_labels(slot, dtype)
        self.label_key[slot] = fingerprint(host)
        return host


def get(shape):
    shape = tuple(int(s) for s in shape)
    s = _sessions.get(shape)
    if s is None:
        # one resident grid at a time keeps HBM use predictable
        close_all()
        s = Session(shape)
        _sessions[shape] = s
    return s


def close_all():
    for s in list(_sessions.values()):
        s.engine.close()
    _sessions.clear()
