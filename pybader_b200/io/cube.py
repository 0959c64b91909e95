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
    file_info = {
        'filename': fn,
        'prefix': prefix,
        'file_type': 'cube',
        'write_function': write,
        'elements': atom_types,
        'voxel_offset': np.array([.5, .5, .5])
    }
    return density, lattice, atoms, file_info


def write(fn, atoms, lattice, density, file_info, prefix=None, suffix='.cube'):
    """Write a cube style charge density (io/cube.py:159-222): same arguments, same file.
    Like the reference, `atoms`, `lattice` and the charge array are converted to Bohr
    units IN PLACE (io/cube.py:183-187)."""
    from ._format import append_block, fortran_lines
    if prefix is not None:
        fn = prefix + fn
    fn += suffix
    fmt = file_info.get('fortran_format', 0)
    charge = density['charge']
    atoms *= ang_to_bohr
    charge *= bohr_to_ang**3
    lattice *= ang_to_bohr
    lattice /= charge.shape
    lattice_width = np.max(np.log10(np.abs(lattice[lattice != 0]))) + 9
    lattice_width = max([int(lattice_width), 9]) + 1
    lattice_prec = 17 - lattice_width
    atoms_width = np.max(np.log10(np.abs(atoms[atoms != 0]))) + 9
    atoms_width = max([int(atoms_width), 9]) + 1
    atoms_prec = 17 - atoms_width
    with open(fn, 'w') as f:
        f.write("Cube File writen in pybader\n")
        f.write(file_info['comment'])
        f.write(f"{atoms.shape[0]:>5}{'  0.0000000'*3}\n")
        for i, lat in enumerate(lattice):
            x, y, z = lat
            f.write(f"{charge.shape[i]:>5}")
            f.write(f" {x:> {10}.{lattice_prec}f} {y:> {10}.{lattice_prec}f} {z:> {10}.{lattice_prec}f}\n")
        for i, atom in enumerate(atoms):
            x, y, z = atom
            f.write(f"{file_info['elements'][i]:>5}")
            f.write('  0.0000000')
            f.write(f" {x:> {10}.{atoms_prec}f} {y:> {10}.{atoms_prec}f} {z:> {10}.{atoms_prec}f}\n")
        if fmt == 2:
            nz = charge.shape[2]
            full = nz // 6 * 6
            for i in range(charge.shape[0]):
                for j in range(charge.shape[1]):
                    row = charge[i, j]
                    if full:
                        f.write(fortran_lines(row[:full].reshape(-1, 6), 5))
                    if full < nz:
                        f.write(fortran_lines(row[full:].reshape(1, -1), 5))
    if fmt != 2:
        append_block(fn, charge, False, charge.shape[2], 6, 5, fmt == 1)
