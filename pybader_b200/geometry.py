"""Host-side geometry the reference derives inside its ``Bader`` object.

When the engine runs underneath the unmodified ``Bader`` class these values
arrive as arguments (``dist_mat``, ``T_grad``, ``voxel_volume``) and this module
is not used.  It exists for callers that have no ``Bader`` object (bench.py,
tests): it performs the *same numpy operations* as the reference so both
engines are fed identical doubles.

    distance_matrix  interface.py:242-259
    voxel_lattice    interface.py:261-265   (np.divide(lattice, shape): column k
                                             is divided by shape[k], kept as is)
    voxel_volume     interface.py:235-240, 267-271
    T_grad           interface.py:285-290
"""
import numpy as np


def voxel_lattice(lattice, shape):
    return np.divide(np.asarray(lattice, dtype=np.float64), shape)


def lattice_volume(lattice):
    lattice = np.asarray(lattice, dtype=np.float64)
    return np.abs(np.dot(lattice[0], np.cross(*lattice[1:])))


def voxel_volume(lattice, shape):
    return lattice_volume(lattice) / np.prod(shape)


def distance_matrix(lattice, shape):
    """3x3x3 table of inverse step lengths; index 2 on an axis means a step of
    -1 (the reference reads it with negative indices, methods.py:110)."""
    vl = voxel_lattice(lattice, shape)
    d = np.zeros((3, 3, 3, 3), dtype=np.float64)
    d[1, :, :] += vl[0]
    d[2, :, :] -= vl[0]
    d[:, 1, :] += vl[1]
    d[:, 2, :] -= vl[1]
    d[:, :, 1] += vl[2]
    d[:, :, 2] -= vl[2]
    d = d**2
    d = np.sum(d, axis=3)
    d[d != 0] = d[d != 0]**-.5
    return d


def T_grad(lattice, shape):
    inv_l = np.linalg.inv(voxel_lattice(lattice, shape))
    return np.matmul(inv_l.T, inv_l)


def maxima_fractional(maxima_idx, shape, voxel_offset_fractional=(0., 0., 0.)):
    """Bader.bader_maxima setter (interface.py:318-324)."""
    m = np.add(maxima_idx, voxel_offset_fractional)
    return np.ascontiguousarray(np.divide(m, shape))
