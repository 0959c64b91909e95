"""Text readers: the decimal -> double conversion on the CPU
against Python's float(), and the GPU readers against fixtures produced by the real
reference readers (tests/golden_io/make_io_golden.py)."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden_io')


@pytest.fixture(scope='module')
def lib():
    from pybader_b200 import build, _lib
    build.build()
    return _lib.load()


def _tokens(rng, n):
    out = []
    for _ in range(n):
        style = rng.integers(0, 6)
        nd = int(rng.integers(1, 20))
        if style == 0:
            nd = 11
        if style == 1:
            nd = 6
        digits = str(rng.integers(1, 10)) + ''.join(str(d) for d in rng.integers(0, 10, nd - 1))
        e = int(rng.integers(-125, 30))
        sg = '-' if rng.random() < 0.25 else ''
        if style == 0:
            out.append(f"{sg}0.{digits}E{e:+03d}")
        elif style == 1:
            out.append(f"{sg}{digits[0]}.{digits[1:]}E{e:+03d}")
        elif style == 2:
            out.append(f"{sg}{digits}")
        elif style == 3:
            k = int(rng.integers(0, len(digits) + 1))
            out.append(f"{sg}{digits[:k]}.{digits[k:]}e{e}")
        elif style == 4:
            out.append(f"{sg}0.{'0' * int(rng.integers(0, 30))}{digits}")
        else:
            out.append(f"+{digits}.E{e}")
    return out


def test_token_conversion_is_correctly_rounded(lib):
    """every token the C conversion accepts equals float(token) bit for bit; what it
    declines (status 2) is left to the host"""
    rng = np.random.default_rng(2024)
    toks = _tokens(rng, 200000)
    toks += ['0', '-0', '0.0', '-0.000E+00', '1', '9007199254740993', '0.1', '1e22', '1e23', '1e-22',
             '1e-23', '123456789012345678', '4.9e-324', '2.2250738585072014e-308', '1.7976931348623157e308',
             '0.30000000000000004', '5e-1', '.5', '5.', '1E5', '8.5E-110', '1.23E-130']
    v = ctypes.c_double(0)
    done = 0
    for t in toks:
        b = t.encode()
        st = lib.bdr_parse_token_host(b, len(b), ctypes.byref(v))
        assert st in (0, 2)
        if st == 0:
            done += 1
            assert np.float64(v.value).tobytes() == np.float64(float(t)).tobytes(), t
    assert done > 0.6 * len(toks)
    for junk in (b'****', b'nan', b'1.0x', b'1e', b'--1', b'.', b'1D5', b'0x10'):
        assert lib.bdr_parse_token_host(junk, len(junk), ctypes.byref(v)) == 2


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(GOLDEN, 'io_golden.npz'))


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['CHGCAR_a', 'CHGCAR_b'])
def test_chgcar_reader_matches_reference(name, golden, capsys):
    from pybader_b200.io import vasp
    d, lat, at, info = vasp.read(os.path.join(GOLDEN, name), charge_flag=True, spin_flag=True)
    for key in ('charge', 'spin'):
        ref = golden[f'{name}_{key}']
        assert d[key].dtype == np.float64 and d[key].flags['C_CONTIGUOUS']
        assert d[key].shape == ref.shape
        assert d[key].tobytes() == ref.tobytes(), f"{name} {key}: not bit-identical"
    np.testing.assert_array_equal(lat, golden[f'{name}_lattice'])
    np.testing.assert_array_equal(at, golden[f'{name}_atoms'])
    np.testing.assert_array_equal(info['element_nums'], golden[f'{name}_element_nums'])
    assert info['file_type'] == 'VASP' and info['spin_flag'] and info['elements'] == ['Si', 'O']
    d2, *_ = vasp.read(os.path.join(GOLDEN, name), charge_flag=True, spin_flag=False)
    assert set(d2) == {'charge'}


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['a.cube', 'b.cube'])
def test_cube_reader_matches_reference(name, golden, capsys):
    from pybader_b200.io import cube
    d, lat, at, info = cube.read(os.path.join(GOLDEN, name))
    ref = golden[f'{name}_charge']
    assert d['charge'].shape == ref.shape and d['charge'].flags['C_CONTIGUOUS']
    assert d['charge'].tobytes() == ref.tobytes()
    np.testing.assert_array_equal(lat, golden[f'{name}_lattice'])
    np.testing.assert_array_equal(at, golden[f'{name}_atoms'])
    np.testing.assert_array_equal(info['elements'], golden[f'{name}_elements'])
    np.testing.assert_array_equal(info['voxel_offset'], [.5, .5, .5])


@pytest.mark.gpu
def test_parse_block_large_and_fallback(tmp_path):
    """a block larger than one 256 MB chunk would be slow here; instead: 3 M tokens with
    junk-free odd formats, a '1D5'-free fallback mix (20-digit mantissas, subnormals),
    both layouts, against numpy's own conversion"""
    from pybader_b200.io._text import OP_DIVIDE, OP_MULTIPLY, parse_block
    rng = np.random.default_rng(99)
    shape = (60, 50, 40)
    n = int(np.prod(shape))
    vals = rng.lognormal(0, 8, n) * rng.choice([1, -1], n)
    toks = np.array(['%.11E' % v for v in vals], dtype=object)
    odd = rng.choice(n, 3000, replace=False)
    toks[odd[:1000]] = ['%.21e' % vals[i] for i in odd[:1000]]        # 22 digits: host fallback
    toks[odd[1000:2000]] = ['%.5e' % (vals[i] * 1e-315) for i in odd[1000:2000]]   # subnormal / zero
    toks[odd[2000:]] = ['%d' % int(vals[i]) for i in odd[2000:]]
    text = ('\n'.join(' '.join(toks[i:i + 7]) for i in range(0, n, 7)) + '\n').encode()
    ref = np.array([float(t) for t in toks])
    a, used = parse_block(text, shape, False, OP_MULTIPLY, 1.25)
    assert used <= len(text)
    assert a.tobytes() == (ref * 1.25).reshape(shape).tobytes()
    b, _ = parse_block(text, shape, True, OP_DIVIDE, 3.7)
    want = np.swapaxes((ref / 3.7).reshape(shape[::-1]), 0, -1).copy()
    assert b.tobytes() == want.tobytes()
    with pytest.raises(ValueError):
        parse_block(text[:len(text) // 2], shape, False)
    with pytest.raises(ValueError):
        parse_block(text.replace(b'E+00', b'*+00', 1), shape, False)


@pytest.mark.gpu
def test_spin_only_read_without_augmentation_block(tmp_path, capsys):
    """a CHG-style file (no augmentation occupancies) whose grid is a multiple of
    the values per line: the second grid line starts exactly where the charge block is
    estimated to end, and the spin-only read must still find it"""
    from pybader_b200.io import vasp
    rng = np.random.default_rng(3)
    nx, ny, nz = 5, 4, 3
    n = nx * ny * nz
    path = os.path.join(str(tmp_path), 'CHG_spin')
    with open(path, 'w') as f:
        f.write("chg fixture\n   1.00000000000000\n")
        f.write("     4.000000    0.000000    0.000000\n     0.000000    5.000000    0.000000\n"
                "     0.000000    0.000000    6.000000\n")
        f.write("   Si\n     1\nDirect\n  0.100000  0.200000  0.300000\n\n")
        for b in range(2):
            f.write(f"   {nx}   {ny}   {nz}\n")
            vals = rng.normal(0, 5, n)
            for i in range(0, n, 5):
                f.write(''.join(' %17.11E' % v for v in vals[i:i + 5]) + '\n')
    both, *_ = vasp.read(path, charge_flag=True, spin_flag=True)
    only, _, _, info = vasp.read(path, charge_flag=False, spin_flag=True)
    assert info['spin_flag'] and set(only) == {'spin'}
    assert only['spin'].tobytes() == both['spin'].tobytes()
    assert not np.array_equal(both['spin'], both['charge'])


@pytest.mark.gpu
def test_fallback_tokens_longer_than_64_bytes():
    """the device hands back at most 64 bytes of a token; the whole token decides"""
    from pybader_b200.io._text import parse_block
    long_ok = '0.' + '123456789' * 8 + 'E+01'          # 78 characters, a valid number
    text = f"1.5 {long_ok} -2.25 4.0\n".encode()
    a, _ = parse_block(text, (1, 2, 2), False)
    assert a.reshape(-1).tolist() == [1.5, float(long_ok), -2.25, 4.0]
    junk = '1.' + '0' * 70 + 'x'
    with pytest.raises(ValueError):
        parse_block(f"1.5 {junk} -2.25 4.0\n".encode(), (1, 2, 2), False)


@pytest.mark.parametrize('name', ['a.cube', 'b.cube'])
def test_cube_header_matches_reference_reader(name, golden):
    """the header half of cube.read needs no GPU: lattice, atoms and elements as the real
    reference reader returned them (tests/golden_io/make_io_golden.py)"""
    from pybader_b200.io import cube
    with open(os.path.join(GOLDEN, name), 'rb') as f:
        hdr = cube._Header(f)
    np.testing.assert_array_equal(hdr.lattice * cube.bohr_to_ang, golden[f'{name}_lattice'])
    np.testing.assert_array_equal(hdr.atoms * cube.bohr_to_ang, golden[f'{name}_atoms'])
    np.testing.assert_array_equal(hdr.atom_types, golden[f'{name}_elements'])
    assert tuple(hdr.grid) == golden[f'{name}_charge'].shape[-3:] or hdr.nval > 1
