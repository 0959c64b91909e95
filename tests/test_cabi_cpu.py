"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports
every symbol include/bader_b200.h declares; host logic; loud failure without a
GPU.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from pybader_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'bader_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bdr_[a-z_0-9]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    from pybader_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string(lib):
    assert lib.bdr_version() >= 100
    assert isinstance(lib.bdr_last_error(), bytes)


def test_create_fails_loudly_without_gpu(lib):
    n = ctypes.c_int(0)
    lib.bdr_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    from pybader_b200.engine import Engine
    from pybader_b200._lib import BaderB200Error
    with pytest.raises(BaderB200Error):
        Engine((8, 8, 8))


def test_bad_arguments(lib):
    h = ctypes.c_void_p()
    assert lib.bdr_create(0, 0, 4, 4, ctypes.byref(h)) != 0
    assert b'empty' in lib.bdr_last_error()
    assert lib.bdr_create(0, 2048, 2048, 2048, ctypes.byref(h)) != 0
    assert b'too large' in lib.bdr_last_error()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, 'pybader_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in text.lower(), \
                    f"{f} mentions the oracle: the product path must not use it"


def test_dtype_calc_matches_reference_table():
    from pybader_b200.utils import dtype_calc
    from oracle.pyoracle import dtype_calc as ref
    for v in (-1, -127, -128, -129, -32767, -32768, -2147483647, -2147483648, -2**40,
              0, 255, 256, 65535, 65536, 4294967295, 4294967296):
        assert dtype_calc(v) == ref(v)
    assert dtype_calc(-127) == 'int8' and dtype_calc(-128) == 'int16'
    assert dtype_calc(-884736) == 'int32'


def test_geometry_matches_reference_formulas():
    from pybader_b200 import geometry as geo
    lat = np.array([[18.0, 0, 0], [4.5, 16.5, 0], [2.4, 3.3, 24.0]])
    shape = (20, 24, 28)
    d = geo.distance_matrix(lat, shape)
    assert d.shape == (3, 3, 3) and d[0, 0, 0] == 0
    vl = geo.voxel_lattice(lat, shape)
    assert np.isclose(d[1, 0, 0], 1 / np.linalg.norm(vl[0]))
    assert np.isclose(d[2, 2, 2], 1 / np.linalg.norm(-vl[0] - vl[1] - vl[2]))
    T = geo.T_grad(lat, shape)
    assert np.allclose(T, T.T)


def test_session_fingerprint_detects_change():
    """residency keys hash the WHOLE array: an edit of any single element changes the
    key (the old strided sample missed most voxels); equal content is
    equal data, whatever buffer holds it"""
    from pybader_b200.session import content_hash, fingerprint
    a = np.zeros((8, 8, 8))
    k = fingerprint(a)
    assert fingerprint(a) == k
    assert content_hash(a)[1] is True
    for idx in np.ndindex(a.shape):
        a[idx] = 1e-300
        assert fingerprint(a) != k, idx
        assert content_hash(a)[1] is False
        a[idx] = 0
    assert fingerprint(a.copy()) == fingerprint(a)
    assert fingerprint(a.astype(np.float32)) != fingerprint(a)
    rng = np.random.default_rng(5)
    b = rng.integers(-1, 100, size=(37, 41, 43)).astype(np.int16)   # ragged: tail bytes, several blocks
    kb = fingerprint(b)
    flat = b.reshape(-1)
    for pos in list(rng.integers(0, flat.size, 200)) + [0, flat.size - 1]:
        old = flat[pos]
        flat[pos] = old + 1
        assert fingerprint(b) != kb, pos
        flat[pos] = old
    assert fingerprint(b) == kb
    # swapping two blocks of the buffer is a different array
    c = np.arange(1 << 16, dtype=np.int64)
    d = np.concatenate([c[1 << 15:], c[:1 << 15]])
    assert fingerprint(c) != fingerprint(d)
    # independent of the thread count
    import ctypes
    from pybader_b200 import _lib
    big = rng.random(3_000_000)
    hs = []
    for nt in (1, 2, 7):
        h, z = ctypes.c_uint64(0), ctypes.c_int(0)
        assert _lib.load().bdr_host_hash(big.ctypes.data, big.nbytes, nt, ctypes.byref(h), ctypes.byref(z)) == 0
        hs.append(h.value)
    assert len(set(hs)) == 1


def test_handlers_dispatch_to_the_sharded_engine_under_torchrun(monkeypatch):
    """WORLD_SIZE > 1 (a torchrun launch) routes every reference-shaped entry point to
    pybader_b200.sharded_handlers; WORLD_SIZE 1 keeps the single-GPU session"""
    from pybader_b200 import sharded_handlers as sh, thread_handlers as th, utils as ut
    monkeypatch.delenv('WORLD_SIZE', raising=False)
    assert not sh.active()
    monkeypatch.setenv('WORLD_SIZE', '1')
    assert not sh.active()
    monkeypatch.setenv('WORLD_SIZE', '4')
    assert sh.active()
    calls = []
    for name in ('bader_calc', 'refine', 'assign_to_atoms', 'surface_distance', 'vacuum_assign',
                 'charge_sum', 'volume_mask'):
        monkeypatch.setattr(sh, name, (lambda n: (lambda *a, **k: calls.append(n) or n))(name))
    sh.refine.last_history = []
    a = np.zeros((2, 2, 2))
    assert th.bader_calc('ongrid', a, a, a, a, 1) == 'bader_calc'
    th.refine('neargrid', ('all', 1), a, a, a, a, 1)
    assert th.assign_to_atoms(a, a, a, a, 1) == 'assign_to_atoms'
    assert th.surface_distance(a, a, a, a, 1) == 'surface_distance'
    assert ut.vacuum_assign(a, a, 0.0, a, 1.0) == 'vacuum_assign'
    assert ut.charge_sum(a, a, 1.0, a, a) == 'charge_sum'
    assert ut.volume_mask(a, a, 0) == 'volume_mask'
    assert calls == ['bader_calc', 'refine', 'assign_to_atoms', 'surface_distance', 'vacuum_assign',
                     'charge_sum', 'volume_mask']
    monkeypatch.setenv('BDR_FORCE_SINGLE', '1')
    assert not sh.active()


def test_geometry_helper_is_bit_exact_against_the_reference_bader():
    """pybader_b200/geometry.py (what tests and bench.py hand to the engine) reproduces
    Bader.distance_matrix / T_grad / voxel_volume (interface.py:242-290) bit for bit: golden
    values stored by tests/golden_call/make_call_golden.py from the real `Bader` object
    (cubic, triclinic and orthorhombic cells)"""
    from pybader_b200 import geometry as geo
    gdir = os.path.join(ROOT, 'tests', 'golden_call')
    shapes = {'c1_default': (96, 96, 96), 'ref_spin': (32, 36, 40), 'c4_slab': (512, 512, 1024)}
    for name, shape in shapes.items():
        with np.load(os.path.join(gdir, name + '.npz')) as z:
            lattice, d, T, dV = z['lattice'], z['distance_matrix'], z['T_grad'], float(z['voxel_volume'])
        np.testing.assert_array_equal(geo.distance_matrix(lattice, shape), d, err_msg=name)
        np.testing.assert_array_equal(geo.T_grad(lattice, shape), T, err_msg=name)
        assert geo.voxel_volume(lattice, shape) == dV, name
