"""Gaussian / CP2K cube reader: the reference's `pybader.io.cube.read`
(io/cube.py:18-156) with the volumetric block converted on the GPU.  Same
arguments, return tuple, unit conversion (Bohr -> Angstrom, values * bohr^-3)."""
import os
from time import time

import numpy as np

from ._text import OP_MULTIPLY, OP_NONE, parse_block

# io/cube.py:13-14
bohr_to_ang = 0.52917721067
ang_to_bohr = 1 / bohr_to_ang


def read(fn, orbitals=0, device=0):
    """Read the charge density from a cube file (io/cube.py:18-156)."""
    t0 = time()
    density = dict()
    prefix, filename = os.path.split(fn)
    prefix = os.path.join(prefix, '')
    with open(fn, 'rb') as f:
        print(f"  Reading {fn} as cube format.")
        _ = f.readline()
        _ = f.readline()
        line = f.readline().split()
        atom_sum = int(line[0])
        origin = np.array(line[1:4], dtype=np.float64)  # noqa: F841 (ignored by the reference too)
        nval = int(line[5]) if len(line) > 4 else 1
        grid = np.zeros(3, dtype=np.int64)
        lattice = np.zeros((3, 3), dtype=np.float64)
        for i in range(3):
            line = f.readline().split()
            grid[i] = line[0]
            lattice[i] = line[1:]
            lattice[i] *= grid[i]
        print(f"  {' x '.join(grid.astype(str))} grid size.")
        atom_types = np.zeros(abs(atom_sum), dtype=np.int64)
        atoms = np.zeros((abs(atom_sum), 3), dtype=np.float64)
        for i in range(abs(atom_sum)):
            line = f.readline().split()
            atom_types[i] = line[0]
            atoms[i] = line[-3:]
        atoms = np.dot(atoms, np.linalg.inv(lattice))
        atoms %= 1
        atoms = np.dot(atoms, lattice)
        dset_ids = None
        if atom_sum < 0:
            line = f.readline().split()
            dset_ids = np.zeros(int(line.pop(0)), dtype=np.int64)
            nval = dset_ids.shape[0]
            count = 0
            while count < nval:
                for m in line:
                    dset_ids[count] = m
                    count += 1
                line = f.readline().split() if count < nval else line
        nx, ny, nz = (int(g) for g in grid)
        buf = np.fromfile(f, dtype=np.uint8)
    # the file runs z (and the nval values of a voxel) fastest: already C order
    # (with one value per voxel the unit conversion, io/cube.py:143, rides along)
    if nval == 1:
        charge, _ = parse_block(buf, (nx, ny, nz), False, OP_MULTIPLY, ang_to_bohr**3, device)
    else:
        charge, _ = parse_block(buf, (nx, ny, nz * nval), False, OP_NONE, 1.0, device)
    del buf
    print(f"  File {fn} closed. ", end='')
    if nval > 1:
        ids = list(dset_ids) if dset_ids is not None else None
        charge = np.swapaxes(charge.reshape(nx, ny, nz, nval), 0, -1)
        if hasattr(orbitals, '__iter__'):
            density['charge'] = np.sum([charge[ids.index(int(m))] for m in orbitals], axis=0)
        elif orbitals < 0:
            density['charge'] = charge
        elif orbitals > 0:
            density['charge'] = charge[ids.index(int(orbitals))].copy()
        elif atom_sum > 0:
            density['charge'] = charge[0].copy()
        else:
            density['charge'] = np.sum(charge, axis=0)
        del charge
    else:
        density['charge'] = charge
    print(f"Time taken: {time() - t0:0.3f}s", end='\n\n')
    lattice *= bohr_to_ang
    atoms *= bohr_to_ang
    if nval > 1:
        density['charge'] *= ang_to_bohr**3
    try:
        from pybader.io.cube import write
    except Exception:                              # noqa: BLE001
        write = None
    file_info = {
        'filename': fn,
        'prefix': prefix,
        'file_type': 'cube',
        'write_function': write,
        'elements': atom_types,
        'voxel_offset': np.array([.5, .5, .5])
    }
    return density, lattice, atoms, file_info
