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


class _Header:
    """Everything a cube file says before its volumetric block (io/cube.py:45-98), pulled
    from a token stream so that the layout of the header lines does not matter."""

    def __init__(self, f):
        f.readline(), f.readline()                      # two comment lines
        first = f.readline().split()
        self.atom_sum = int(first[0])                   # negative: a data-set id line follows the atoms
        self.nval = int(first[5]) if len(first) > 4 else 1   # (the origin, first[1:4], is ignored by the reference)
        n_atoms = abs(self.atom_sum)
        axes = np.array([f.readline().split() for _ in range(3)], dtype=np.float64)
        self.grid = axes[:, 0].astype(np.int64)
        self.lattice = axes[:, 1:4] * axes[:, :1]       # voxel vectors times the voxel counts
        rows = [f.readline().split() for _ in range(n_atoms)]
        self.atom_types = np.array([r[0] for r in rows], dtype=np.int64).reshape(n_atoms)
        pos = np.array([r[-3:] for r in rows], dtype=np.float64).reshape(n_atoms, 3)
        frac = np.dot(pos, np.linalg.inv(self.lattice))
        frac %= 1
        self.atoms = np.dot(frac, self.lattice)         # wrapped into the cell, still in Bohr
        self.dset_ids = None
        if self.atom_sum < 0:
            ids = f.readline().split()
            self.nval = int(ids.pop(0))
            while len(ids) < self.nval:                 # the id list may run over several lines
                ids += f.readline().split()
            self.dset_ids = [int(m) for m in ids[:self.nval]]


def _select_orbitals(values, hdr, orbitals):
    """what `orbitals` asks of a file with several values per voxel (io/cube.py:114-141);
    values is [nval][nx][ny][nz]"""
    if hasattr(orbitals, '__iter__'):
        return np.sum([values[hdr.dset_ids.index(int(m))] for m in orbitals], axis=0)
    if orbitals < 0:
        return values
    if orbitals > 0:
        return values[hdr.dset_ids.index(int(orbitals))].copy()
    if hdr.atom_sum > 0:
        return values[0].copy()
    return np.sum(values, axis=0)


def read(fn, orbitals=0, device=0):
    """Read the charge density from a cube file (io/cube.py:18-156): same arguments, same
    return tuple, same units (Bohr -> Angstrom, values * bohr^-3)."""
    t0 = time()
    prefix = os.path.join(os.path.split(fn)[0], '')
    with open(fn, 'rb') as f:
        print(f"  Reading {fn} as cube format.")
        hdr = _Header(f)
        print(f"  {' x '.join(hdr.grid.astype(str))} grid size.")
        block = np.fromfile(f, dtype=np.uint8)
    nx, ny, nz = (int(g) for g in hdr.grid)
    # the file runs z (and the nval values of a voxel) fastest: already C order; with one
    # value per voxel the unit conversion (io/cube.py:143) rides along in the conversion
    if hdr.nval == 1:
        charge, _ = parse_block(block, (nx, ny, nz), False, OP_MULTIPLY, ang_to_bohr**3, device)
    else:
        raw, _ = parse_block(block, (nx, ny, nz * hdr.nval), False, OP_NONE, 1.0, device)
        charge = _select_orbitals(np.swapaxes(raw.reshape(nx, ny, nz, hdr.nval), 0, -1), hdr, orbitals)
        charge *= ang_to_bohr**3
    del block
    print(f"  File {fn} closed. ", end='')
    print(f"Time taken: {time() - t0:0.3f}s", end='\n\n')
    file_info = dict(filename=fn, prefix=prefix, file_type='cube', write_function=write,
                     elements=hdr.atom_types, voxel_offset=np.array([.5, .5, .5]))
    return {'charge': charge}, hdr.lattice * bohr_to_ang, hdr.atoms * bohr_to_ang, file_info


def _fixed(lead, row, prec):
    """one header line: `lead`, then three coordinates, each 10 wide with `prec` decimals
    (as many as the widest entry of the table leaves room for, io/cube.py:190-195)"""
    return lead + ''.join(f" {v:> 10.{prec}f}" for v in row) + '\n'


def _precision(table):
    width = max(int(np.max(np.log10(np.abs(table[table != 0]))) + 9), 9) + 1
    return 17 - width


def write(fn, atoms, lattice, density, file_info, prefix=None, suffix='.cube'):
    """Write a cube style charge density (io/cube.py:159-222): same arguments, same file.
    Like the reference, `atoms`, `lattice` and the charge array are converted to Bohr
    units IN PLACE (io/cube.py:183-187)."""
    from ._format import append_block, fortran_lines
    path = (prefix or '') + fn + suffix
    fmt = file_info.get('fortran_format', 0)
    charge = density['charge']
    atoms *= ang_to_bohr
    charge *= bohr_to_ang**3
    lattice *= ang_to_bohr
    lattice /= charge.shape
    lat_prec, atom_prec = _precision(lattice), _precision(atoms)
    head = ["Cube File writen in pybader\n", file_info['comment'],
            f"{atoms.shape[0]:>5}{'  0.0000000' * 3}\n"]
    head += [_fixed(f"{n:>5}", vec, lat_prec) for n, vec in zip(charge.shape, lattice)]
    head += [_fixed(f"{el:>5}  0.0000000", xyz, atom_prec) for el, xyz in zip(file_info['elements'], atoms)]
    with open(path, 'w') as f:
        f.writelines(head)
        if fmt == 2:
            # utils.fortran_format: six values per line, a shorter last line per z row
            nz = charge.shape[2]
            full = nz // 6 * 6
            for row in charge.reshape(-1, nz):
                if full:
                    f.write(fortran_lines(row[:full].reshape(-1, 6), 5))
                if full < nz:
                    f.write(fortran_lines(row[full:].reshape(1, -1), 5))
    if fmt != 2:
        append_block(path, charge, False, charge.shape[2], 6, 5, fmt == 1)
