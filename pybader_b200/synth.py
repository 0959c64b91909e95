"""Synthetic Gaussian-superposition densities.

rho(r) = sum_a A_a sum_{27 images T} exp(-|r - R_a - T|^2 / (2 sigma_a^2)),
grid point (i,j,k) at fractional (i/nx, j/ny, k/nz), array layout [x][y][z].

The host generator below serves the parity-sized cases; the big bench cases
are generated on the device (csrc, ``bdr_synth_*``) and read back, so that in
every comparison both engines consume the very same bytes.
"""
import numpy as np


def gaussian_density(shape, lattice, frac_atoms, amps, sigmas, chunk=8):
    shape = tuple(int(s) for s in shape)
    lattice = np.asarray(lattice, dtype=np.float64)
    frac_atoms = np.asarray(frac_atoms, dtype=np.float64).reshape(-1, 3)
    amps = np.broadcast_to(np.asarray(amps, dtype=np.float64), (len(frac_atoms),))
    sigmas = np.broadcast_to(np.asarray(sigmas, dtype=np.float64), (len(frac_atoms),))
    fy = np.arange(shape[1]) / shape[1]
    fz = np.arange(shape[2]) / shape[2]
    images = np.array([(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1)
                       for c in (-1, 0, 1)], dtype=np.float64)
    rho = np.zeros(shape, dtype=np.float64)
    for x0 in range(0, shape[0], chunk):
        fx = np.arange(x0, min(x0 + chunk, shape[0])) / shape[0]
        f = np.stack(np.meshgrid(fx, fy, fz, indexing='ij'), axis=-1)
        for atom, amp, sig in zip(frac_atoms, amps, sigmas):
            for im in images:
                d = (f - (atom + im)) @ lattice
                r2 = np.einsum('...k,...k->...', d, d)
                rho[x0:x0 + len(fx)] += amp * np.exp(-r2 / (2.0 * sig * sig))
    return rho


def case_c1(n=96):
    """BASELINE config 1: 3-atom cubic cell, a = 6.0 A."""
    lattice = np.eye(3) * 6.0
    frac = np.array([[0.25, 0.25, 0.25], [0.66, 0.5, 0.5], [0.5, 0.75, 0.33]])
    amps = np.array([1.0, 0.8, 1.3])
    return dict(shape=(n, n, n), lattice=lattice, frac_atoms=frac, amps=amps,
                sigmas=np.full(3, 0.5))


def case_rocksalt(n=256, cells=4, offset=0.13, a=11.28):
    """BASELINE config 2: rocksalt-like cells^3 sites on an n^3 grid."""
    idx = np.array([(i, j, k) for i in range(cells) for j in range(cells)
                    for k in range(cells)], dtype=np.float64)
    frac = (idx + offset) / cells
    even = (idx.sum(axis=1).astype(int) % 2) == 0
    amps = np.where(even, 1.0, 2.2)
    sig = np.where(even, 0.35, 0.55) * (a / 11.28) * (4.0 / cells)
    return dict(shape=(n, n, n), lattice=np.eye(3) * a, frac_atoms=frac, amps=amps,
                sigmas=sig)


def case_triclinic(shape=(360, 360, 480), n_atoms=128, seed=1234):
    """BASELINE config 3: triclinic cell with vacuum around a slab of atoms."""
    lattice = np.array([[18.0, 0, 0], [4.5, 16.5, 0], [2.4, 3.3, 24.0]])
    rng = np.random.default_rng(seed)
    frac = np.empty((n_atoms, 3))
    frac[:, 0] = rng.uniform(0.1, 0.9, n_atoms)
    frac[:, 1] = rng.uniform(0.1, 0.9, n_atoms)
    frac[:, 2] = rng.uniform(0.25, 0.75, n_atoms)
    amps = rng.uniform(0.8, 2.5, n_atoms)
    sig = rng.uniform(0.35, 0.6, n_atoms)
    return dict(shape=tuple(shape), lattice=lattice, frac_atoms=frac, amps=amps,
                sigmas=sig)


def case_slab(shape=(512, 512, 1024), n_atoms=64, seed=4321):
    """BASELINE config 4: orthorhombic slab with vacuum, plus a spin density
    built from the same Gaussians with weights +-0.3."""
    lattice = np.diag([12.0, 12.0, 24.0])
    rng = np.random.default_rng(seed)
    frac = np.empty((n_atoms, 3))
    frac[:, 0] = rng.uniform(0.0, 1.0, n_atoms)
    frac[:, 1] = rng.uniform(0.0, 1.0, n_atoms)
    frac[:, 2] = rng.uniform(0.3, 0.7, n_atoms)
    amps = rng.uniform(0.8, 2.5, n_atoms)
    sig = rng.uniform(0.35, 0.6, n_atoms)
    spin_w = np.where(rng.uniform(size=n_atoms) < 0.5, -0.3, 0.3)
    return dict(shape=tuple(shape), lattice=lattice, frac_atoms=frac, amps=amps,
                sigmas=sig, spin_weights=spin_w)


def case_lattice_sites(shape, cells, a, seed=2048, jitter=0.13, sigma_frac=0.11):
    """BASELINE config 5 family: cells[0]*cells[1]*cells[2] atoms on a jittered
    simple lattice in an orthorhombic cell (separable -> device generator)."""
    cells = tuple(int(c) for c in cells)
    rng = np.random.default_rng(seed)
    idx = np.array([(i, j, k) for i in range(cells[0]) for j in range(cells[1])
                    for k in range(cells[2])], dtype=np.float64)
    frac = (idx + 0.5 + rng.uniform(-jitter, jitter, idx.shape)) / np.array(cells)
    n = len(frac)
    amps = rng.uniform(0.8, 2.5, n)
    spacing = min(a[i] / cells[i] for i in range(3))
    sig = rng.uniform(0.8, 1.2, n) * sigma_frac * spacing
    return dict(shape=tuple(shape), lattice=np.diag(np.asarray(a, dtype=np.float64)),
                frac_atoms=frac, amps=amps, sigmas=sig)


def make(case, spin=False):
    rho = gaussian_density(case['shape'], case['lattice'], case['frac_atoms'],
                           case['amps'], case['sigmas'])
    atoms_cart = case['frac_atoms'] @ case['lattice']
    if spin:
        w = case.get('spin_weights')
        if w is None:
            w = np.where(np.arange(len(case['amps'])) % 2 == 0, 0.3, -0.3)
        s = gaussian_density(case['shape'], case['lattice'], case['frac_atoms'],
                             case['amps'] * w, case['sigmas'])
        return rho, s, atoms_cart
    return rho, atoms_cart


def separable_tables(case):
    """1-D factor tables for an orthorhombic cell: rho[i,j,k] =
    sum_a tx[a,i] * ty[a,j] * tz[a,k] (amplitude folded into tx; the 27-image
    sum factorises into three 3-image sums)."""
    lattice = np.asarray(case['lattice'], dtype=np.float64)
    assert np.count_nonzero(lattice - np.diag(np.diag(lattice))) == 0, "orthorhombic only"
    frac = np.asarray(case['frac_atoms'], dtype=np.float64).reshape(-1, 3)
    amps = np.broadcast_to(np.asarray(case['amps'], dtype=np.float64), (len(frac),))
    sig = np.broadcast_to(np.asarray(case['sigmas'], dtype=np.float64), (len(frac),))
    tabs = []
    for ax in range(3):
        n = case['shape'][ax]
        f = np.arange(n) / n
        t = np.zeros((len(frac), n))
        for m in (-1, 0, 1):
            d = (f[None, :] - frac[:, ax:ax + 1] - m) * lattice[ax, ax]
            t += np.exp(-d * d / (2.0 * sig[:, None] ** 2))
        tabs.append(t)
    tabs[0] = tabs[0] * amps[:, None]
    return tuple(np.ascontiguousarray(t) for t in tabs)
