"""CHGCAR / CHG reader: the reference's `pybader.io.vasp.read` (io/vasp.py:15-164)
with the density blocks converted on the GPU.  Same arguments, same return
tuple, same arithmetic (tokens -> float64, / cell volume, [x][y][z] C order)."""
import os
from time import time

import numpy as np

from ._text import OP_DIVIDE, parse_block


def _read_block(f, n_values, est_bytes):
    """bytes of a block of n_values tokens starting at the file position: the
    reference's fixed-line-length estimate plus slack, grown if tokens are missing"""
    start = f.tell()
    f.seek(0, 2)
    end = f.tell()
    f.seek(start)
    want = min(end - start, est_bytes + 4096)
    return start, end, np.fromfile(f, dtype=np.uint8, count=want)


def read(fn, charge_flag=True, spin_flag=False, buffer_size=64, device=0):
    """Read the charge and/or spin density from a VASP CHGCAR (io/vasp.py:15-164).

    return: density dict, lattice (rows), atoms (Cartesian), file_info"""
    t0 = time()
    density = dict()
    prefix, filename = os.path.split(fn)
    prefix = os.path.join(prefix, '')
    with open(fn, 'rb') as f:
        print(f"  Reading {fn} as CHGCAR format.")
        _ = f.readline()                                   # comment line of the POSCAR
        scale = np.array(f.readline().split(), dtype=np.float64)
        lattice = np.zeros((3, 3), dtype=np.float64)
        for i in range(3):
            lattice[i] = f.readline().split()
        atom_types = [t.decode() for t in f.readline().split()]
        try:
            atom_nums = np.array(atom_types, dtype=np.int64)   # no symbols line
            atom_types = None
        except ValueError:
            atom_nums = np.array(f.readline().split(), dtype=np.int64)
        atom_sum = int(atom_nums.sum())
        coord_system = f.readline().lstrip().lower()
        atoms = np.zeros((atom_sum, 3), dtype=np.float64)
        for i in range(atom_sum):
            atoms[i] = f.readline().split()
        if coord_system[:1] == b'd':
            atoms %= 1
        else:
            atoms = np.dot(atoms, np.linalg.inv(lattice))
            atoms %= 1
        _ = f.readline()
        grid_line = f.readline()
        grid = np.array(grid_line.split(), dtype=np.int64)
        grid_pts = int(np.prod(grid))
        print(f"  {' x '.join(grid.astype(str))} grid size.")
        charge_pos = f.tell()
        first = f.readline()
        per_line = max(len(first.split()), 1)
        line_len = len(first)
        est = (grid_pts // per_line + 1) * line_len
        # lattice scaling and the cell volume come first here: the division is fused
        # into the conversion (io/vasp.py:140-149 does it afterwards; same operands)
        if scale.shape[0] == 1:
            lattice *= scale[0]
        else:
            for i in range(3):
                lattice[i] *= scale[i]
        lattice_vol = np.dot(lattice[0], np.cross(*lattice[1:]))
        shape = tuple(int(g) for g in grid)

        def block(pos):
            f.seek(pos)
            start, end, buf = _read_block(f, grid_pts, est)
            try:
                arr, used = parse_block(buf, shape, True, OP_DIVIDE, float(lattice_vol), device)
            except ValueError:
                if start + buf.size >= end:
                    raise
                f.seek(start)                              # variable-width lines: take the rest
                buf = np.fromfile(f, dtype=np.uint8)
                arr, used = parse_block(buf, shape, True, OP_DIVIDE, float(lattice_vol), device)
            return arr, start + used

        charge_end = None
        if charge_flag:
            density['charge'], charge_end = block(charge_pos)
        if spin_flag:
            # the spin block follows the augmentation occupancies, behind a second
            # copy of the grid line (io/vasp.py:106-123 looks for it from mid-file)
            # The first piece of the tail is the rest of a line in both cases -- what is left
            # of the last charge line, or (spin only) a line inside the charge block three
            # lines before its estimated end, so that a grid line that starts exactly at the
            # estimate (no augmentation block, grid_pts a multiple of the values per line) is
            # a whole line of the tail and not mistaken for a partial one.
            f.seek(charge_end if charge_end is not None
                   else max(charge_pos, charge_pos + est - 4 * line_len))
            rest_pos = f.tell()
            tail = f.read()
            key = grid_line.strip()
            hit, at = -1, tail.find(b'\n') + 1
            while at > 0:
                nl = tail.find(b'\n', at)
                if nl < 0:
                    break
                if tail[at:nl].strip() == key:
                    hit = nl + 1
                    break
                at = nl + 1
            del tail
            if hit < 0:
                print(f"  No spin density in {fn}")
                spin_flag = False
            else:
                density['spin'], _ = block(rest_pos + hit)
        print(f"  File {fn} closed. ", end='')
    atoms = np.dot(atoms, lattice)
    print(f"Time taken: {time() - t0:0.3f}s", end='\n\n')
    file_info = {
        'filename': filename,
        'prefix': prefix,
        'file_type': 'VASP',
        'buffer_size': buffer_size,
        'write_function': write,
        'element_nums': atom_nums,
        'charge_flag': charge_flag,
        'spin_flag': spin_flag,
        'voxel_offset': np.zeros(3)
    }
    if atom_types is not None:
        file_info['elements'] = atom_types
    return density, lattice, atoms, file_info


def write(fn, atoms, lattice, density, file_info, prefix='', suffix='-CHGCAR'):
    """Write a VASP style charge density (io/vasp.py:167-258): same arguments, same file.

    Like the reference, the density arrays are scaled by the cell volume IN PLACE
    (io/vasp.py:188-190).  The number blocks are formatted natively (`bdr_format_grid`)
    for fortran_format 0 and 1; 2 goes through the numpy restatement of
    utils.fortran_format.  One deliberate difference: a grid whose size is a multiple of
    5 is written correctly (the reference fails on `charge[:-0]`, io/vasp.py:202-203)."""
    from ._format import append_block, fortran_lines
    fn = prefix + fn + suffix
    fmt = file_info.get('fortran_format', 0)
    lattice_vol = np.dot(lattice[0], np.cross(*lattice[1:]))
    for key in density:
        density[key] *= lattice_vol
    blocks = []
    if file_info['charge_flag']:
        blocks.append(density.get('charge'))
    if file_info['spin_flag']:
        blocks.append(density.get('spin'))
    shape = blocks[0].shape
    lattice_width = np.max(np.log10(np.abs(lattice[lattice != 0]))) + 9
    lattice_width = max([int(lattice_width), 9]) + 1
    lattice_prec = 17 - lattice_width
    atoms_width = np.max(np.log10(np.abs(atoms))).astype(int) + 9
    atoms_width = max([atoms_width, 9]) + 1
    atoms_prec = 17 - atoms_width
    with open(fn, 'w') as f:
        f.write(file_info['comment'])
        f.write(f"{1:0< 10.7f}\n")
        for x, y, z in lattice:
            f.write(f" {x:> {10}.{lattice_prec}f} {y:> {10}.{lattice_prec}f} {z:> {10}.{lattice_prec}f}\n")
        if file_info.get('elements', None) is not None:
            f.write('  '.join(file_info['elements']) + '\n')
        f.write('  '.join(file_info['element_nums'].astype(str)) + '\n')
        f.write('Cartesian\n')
        for x, y, z in atoms:
            f.write(f" {x:> {10}.{atoms_prec}f} {y:> {10}.{atoms_prec}f} {z:> {10}.{atoms_prec}f}\n")
        f.write('\n')
    x, y, z = shape
    for block in blocks:
        with open(fn, 'a') as f:
            f.write(f" {x:>5} {y:>5} {z:>5}\n")
        if fmt == 2:
            flat = np.swapaxes(block, 0, -1).flatten()
            full = flat.size // 5 * 5
            with open(fn, 'a') as f:
                if full:
                    f.write(fortran_lines(flat[:full].reshape(-1, 5), 11))
                if full < flat.size:
                    f.write(fortran_lines(flat[full:].reshape(1, -1), 11))
        else:
            # file order is x fastest: one numpy transpose (as the reference does), then a single row
            flat = np.ascontiguousarray(np.swapaxes(block, 0, -1)).reshape(1, 1, -1)
            append_block(fn, flat, False, flat.size, 5, 11, fmt == 1)
