"""Writers (second half): files written by pybader_b200.io.vasp.write /
cube.write are byte-identical to the files the REAL reference writers produced for the same
inputs (tests/golden_io/written, make_write_golden.py), for all three number formats.
Host code only: runs without a GPU."""
import filecmp
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden_io', 'written')


def _inputs(seed, shape):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'make_write_golden', os.path.join(ROOT, 'tests', 'golden_io', 'make_write_golden.py'))
    src = open(spec.origin).read()
    ns = {}
    # only the input generator of the golden script (its imports need the reference)
    start = src.index('def inputs(')
    end = src.index("if __name__ == '__main__':")
    exec('import numpy as np\n' + src[start:end], ns)
    return ns['inputs'](seed, shape)


@pytest.mark.parametrize('fmt', [0, 1, 2])
def test_chgcar_writer_byte_identical(fmt, tmp_path):
    from pybader_b200 import build
    build.build()
    from pybader_b200.io import vasp
    charge, spin, lattice, atoms = _inputs(3 + fmt, (7, 6, 11))
    before = charge.copy()
    info = {'comment': 'golden\n', 'fortran_format': fmt, 'charge_flag': True, 'spin_flag': True,
            'element_nums': np.array([1, 2]), 'elements': ['Si', 'O'], 'buffer_size': 8}
    vasp.write(f'ref{fmt}', atoms, lattice, {'charge': charge, 'spin': spin}, info,
               prefix=os.path.join(str(tmp_path), ''))
    assert filecmp.cmp(os.path.join(str(tmp_path), f'ref{fmt}-CHGCAR'),
                       os.path.join(GOLD, f'ref{fmt}-CHGCAR'), shallow=False)
    # the reference scales the caller's arrays by the cell volume in place; so do we
    vol = np.dot(lattice[0], np.cross(*lattice[1:]))
    np.testing.assert_array_equal(charge, before * vol)


@pytest.mark.parametrize('fmt', [0, 1, 2])
def test_cube_writer_byte_identical(fmt, tmp_path):
    from pybader_b200 import build
    build.build()
    from pybader_b200.io import cube
    charge, spin, lattice, atoms = _inputs(13 + fmt, (4, 5, 13))
    info = {'comment': 'golden cube\n', 'fortran_format': fmt, 'elements': np.array([14, 8, 8])}
    cube.write(f'ref{fmt}', atoms, lattice, {'charge': charge}, info, prefix=os.path.join(str(tmp_path), ''))
    assert filecmp.cmp(os.path.join(str(tmp_path), f'ref{fmt}.cube'),
                       os.path.join(GOLD, f'ref{fmt}.cube'), shallow=False)


def test_chgcar_writer_grid_multiple_of_five(tmp_path):
    """the reference fails when the grid size is a multiple of 5 (`charge[:-0]`,
    io/vasp.py:202-203); this writer writes the file, and the GPU-free part of the reader's
    contract holds for it: 5 values per line, every line the same length"""
    from pybader_b200.io import vasp
    rng = np.random.default_rng(0)
    charge = rng.random((5, 4, 3)) + 0.5
    info = {'comment': 'c\n', 'fortran_format': 0, 'charge_flag': True, 'spin_flag': False,
            'element_nums': np.array([1])}
    vasp.write('m5', np.array([[0.1, 0.2, 0.3]]), np.eye(3) * 3.0, {'charge': charge}, info,
               prefix=os.path.join(str(tmp_path), ''))
    lines = open(os.path.join(str(tmp_path), 'm5-CHGCAR')).read().split('\n')
    body = lines[lines.index('     5     4     3') + 1:]
    body = [ln for ln in body if ln]
    assert len(body) == 12 and all(len(ln.split()) == 5 for ln in body)
    assert len({len(ln) for ln in body}) == 1


@pytest.mark.parametrize('prec,sign_space', [(11, False), (5, False), (11, True), (3, False)])
def test_native_formatter_equals_python_format(prec, sign_space, tmp_path):
    """bdr_format_grid against Python's own formatting (utils.python_format is
    ' {:.{prec}E}' per value), over magnitudes that take the exact path and ones that fall
    back to printf (huge, tiny, zero, negative zero, few-bit mantissas = exact decimal ties)"""
    from pybader_b200.io._format import append_block
    rng = np.random.default_rng(prec + 7 * sign_space)
    n = 6 * 7 * 50
    v = rng.lognormal(0, 6, n) * rng.choice([1, -1], n)
    v[:200] = np.ldexp(1.0 + rng.integers(0, 16, 200) / 16.0, rng.integers(-40, 40, 200))
    v[200:260] = rng.lognormal(0, 3, 60) * 1e15
    v[260:320] = rng.lognormal(0, 3, 60) * 1e-120
    v[320:324] = [0.0, -0.0, 5e-324, 1.7976931348623157e308]
    a = v.reshape(6, 7, 50)
    path = os.path.join(str(tmp_path), 'block.txt')
    append_block(path, a, False, 50, 6, prec, sign_space)
    align = ' ' if sign_space else ''
    want = []
    for row in a.reshape(-1, 50):
        for k in range(0, 50, 6):
            want.append(''.join(f' {x:{align}.{prec}E}' for x in row[k:k + 6]) + '\n')
    assert open(path).read() == ''.join(want)
