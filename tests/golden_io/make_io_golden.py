#!/usr/bin/env python
"""Golden fixtures for the text readers: small CHGCAR (charge + augmentation
occupancies + spin block, with tiny / negative / zero values) and cube files,
read by the REAL reference readers (pybader.io.vasp.read / pybader.io.cube.read,
imported from /root/reference) and stored next to the text files.

    HOME=/tmp/x PYTHONPATH=/root/reference python tests/golden_io/make_io_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
home = os.environ.setdefault('HOME', '/tmp/pybader_home')
cfg = os.path.join(home, '.config', 'bader')
os.makedirs(cfg, exist_ok=True)
if not os.path.exists(os.path.join(cfg, 'config.ini')):
    open(os.path.join(cfg, 'config.ini'), 'w').write(
        "[DEFAULT]\nmethod = neargrid\nrefine_method = neargrid\nvacuum_tol = None\n"
        "refine_mode = ('changed', 2)\nbader_volume_tol = 0.001\nexport_mode = None\nprefix = ''\n"
        "output = pickle\nthreads = 1\nfortran_format = 0\nspeed_flag = False\nspin_flag = False\n"
        "[speed]\nmethod = ongrid\nrefine_method = neargrid\nrefine_mode = ('changed', 3)\nspeed_flag = True\n")
os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(home, 'nc'))
sys.path.insert(0, '/root/reference')
from pybader.io import cube, vasp  # noqa: E402

rng = np.random.default_rng(7)


def fortran_e(v, width=18, digits=11):
    """VASP's 0.dddddddddddE+xx"""
    if v == 0:
        return ' ' + '0.' + '0' * digits + 'E+00'
    s = '%.*E' % (digits - 1, v)
    mant, ex = s.split('E')
    sign = '-' if mant.startswith('-') else ' '
    d = mant.lstrip('-').replace('.', '')
    return f"{sign}0.{d}E{int(ex) + 1:+03d}"


def write_chgcar(path, grid, per_line, spin):
    nx, ny, nz = grid
    n = nx * ny * nz
    with open(path, 'w') as f:
        f.write("golden fixture\n   1.10000000000000\n")
        f.write("     4.000000    0.100000    0.000000\n     0.300000    5.000000    0.200000\n"
                "     0.000000    0.400000    6.000000\n")
        f.write("   Si   O\n     1     2\nDirect\n")
        f.write("  0.100000  0.200000  0.300000\n  0.600000  0.700000  0.800000\n  1.250000 -0.100000  0.500000\n")
        f.write("\n")
        blocks = 2 if spin else 1
        for b in range(blocks):
            f.write(f"   {nx}   {ny}   {nz}\n")
            vals = rng.lognormal(0, 3, n) * rng.choice([1, 1, 1, -1], n)
            vals[rng.random(n) < 0.05] = 0.0
            vals[rng.random(n) < 0.05] *= 1e-30          # beyond the one-multiply fast path
            vals[rng.random(n) < 0.02] *= 1e-70
            toks = [fortran_e(v) for v in vals]
            for i in range(0, n, per_line):
                f.write(''.join('%19s' % t.strip() for t in toks[i:i + per_line]) + '\n')
            if b == 0:
                f.write("augmentation occupancies   1  4\n  0.1234567E+00 -0.7654321E-01  0.1111111E+01  0.0000000E+00\n")
                f.write("augmentation occupancies   2  3\n  0.2222222E+00  0.3333333E+00 -0.4444444E+00\n")


def write_cube(path, grid):
    nx, ny, nz = grid
    with open(path, 'w') as f:
        f.write("golden cube\ncomment\n")
        f.write(f"    2    0.000000    0.000000    0.000000\n")
        f.write(f"  {nx:4d}    0.400000    0.010000    0.000000\n")
        f.write(f"  {ny:4d}    0.000000    0.500000    0.020000\n")
        f.write(f"  {nz:4d}    0.030000    0.000000    0.600000\n")
        f.write("    8    8.000000    1.000000    1.500000    2.000000\n")
        f.write("    1    1.000000    3.500000   -0.500000    6.100000\n")
        for x in range(nx):
            for y in range(ny):
                vals = rng.lognormal(-3, 4, nz)
                vals[rng.random(nz) < 0.1] *= 1e-40
                line = []
                for k, v in enumerate(vals):
                    line.append(' %12.5E' % v)
                    if (k + 1) % 6 == 0 or k == nz - 1:
                        f.write(''.join(line) + '\n')
                        line = []


out = {}
with contextlib.redirect_stdout(io.StringIO()):
    for name, grid, per_line in (('CHGCAR_a', (12, 10, 8), 5), ('CHGCAR_b', (7, 9, 11), 10)):
        p = os.path.join(HERE, name)
        write_chgcar(p, grid, per_line, spin=True)
        d, lat, at, info = vasp.read(p, charge_flag=True, spin_flag=True)
        out[name + '_charge'], out[name + '_spin'] = d['charge'], d['spin']
        out[name + '_lattice'], out[name + '_atoms'] = lat, at
        out[name + '_element_nums'] = info['element_nums']
    for name, grid in (('a.cube', (6, 5, 13)), ('b.cube', (4, 4, 6))):
        p = os.path.join(HERE, name)
        write_cube(p, grid)
        d, lat, at, info = cube.read(p)
        out[name + '_charge'], out[name + '_lattice'], out[name + '_atoms'] = d['charge'], lat, at
        out[name + '_elements'] = info['elements']
np.savez_compressed(os.path.join(HERE, 'io_golden.npz'), **out)
print({k: v.shape for k, v in out.items()})
