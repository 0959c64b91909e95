"""Generate the golden fixtures in this directory by RUNNING THE REAL REFERENCE.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own,
so these fixtures -- inputs together with the outputs of the unmodified numba
kernels and thread handlers -- are the pin for oracle/ and, through it, for the
CUDA path.  Inputs are stored verbatim (numpy's exp is not guaranteed to be
bit-identical across CPUs).  Nothing here is imported by the product.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CONFIG_INI = """[DEFAULT]
method = neargrid
refine_method = neargrid
vacuum_tol = None
refine_mode = ('changed', 2)
bader_volume_tol = 0.001
export_mode = None
prefix = ''
output = pickle
threads = 1
fortran_format = 0
speed_flag = False
spin_flag = False

[speed]
method = ongrid
refine_method = neargrid
refine_mode = ('changed', 3)
speed_flag = True
"""


def import_reference(ref_root='/root/reference', scratch='/tmp/pybader_ref_home'):
    """SURVEY.md section 8c recipe: scratch HOME with a pre-seeded config.ini (values of
    entry_points.py:326-345), writable numba cache, reference on sys.path."""
    cfg = os.path.join(scratch, '.config', 'bader')
    os.makedirs(cfg, exist_ok=True)
    ini = os.path.join(cfg, 'config.ini')
    if not os.path.exists(ini):
        with open(ini, 'w') as f:
            f.write(CONFIG_INI)
    os.environ['HOME'] = scratch
    os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(scratch, 'numba_cache'))
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import pybader  # noqa: F401
    from pybader import methods, refinement, thread_handlers, utils
    from pybader.interface import Bader
    return dict(methods=methods, refinement=refinement, th=thread_handlers,
                utils=utils, Bader=Bader)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def run_reference(ref, rho, lattice, atoms_cart, vacuum_tol=None, spin=None,
                  voxel_offset=(0., 0., 0.)):
    """All stage outputs of the unmodified reference for one input."""
    from pybader_b200 import geometry as geo
    th, utils, refinement = ref['th'], ref['utils'], ref['refinement']
    shape = rho.shape
    dist = geo.distance_matrix(lattice, shape)
    T = geo.T_grad(lattice, shape)
    dV = geo.voxel_volume(lattice, shape)
    out = dict(rho=rho, lattice=np.asarray(lattice, dtype=np.float64),
               atoms=np.asarray(atoms_cart, dtype=np.float64), dist_mat=dist, T_grad=T,
               voxel_volume=np.float64(dV), voxel_offset=np.asarray(voxel_offset, float),
               vacuum_tol=np.float64(np.nan if vacuum_tol is None else vacuum_tol))
    if spin is not None:
        out['spin'] = spin

    def fresh():
        v = np.zeros(shape, dtype=utils.dtype_calc(-int(np.prod(shape))))
        if vacuum_tol is not None:
            v, q, vol = utils.vacuum_assign(rho, v, np.float64(vacuum_tol), rho, dV)
            out['vacuum_charge'], out['vacuum_volume'] = np.float64(q), np.float64(vol)
        return v

    with quiet():
        # --- ongrid: labels, maxima (bit-exact target) ------------------------
        mx, vo = th.bader_calc('ongrid', rho, fresh(), dist, T, 1)
        out['ongrid_maxima'], out['ongrid_volumes'] = mx, vo.copy()
        # --- edge_find on the ongrid labels ------------------------------------
        known = np.zeros(shape, dtype=np.int8)
        out['ongrid_edges'] = np.int64(refinement.edge_find(known, rho, vo))
        out['ongrid_known'] = known.copy()
        # --- one Jacobi trace iteration on those edges -------------------------
        v1 = vo.copy()
        k1, ch = refinement.neargrid(known.copy(), known.copy(), rho, v1,
                                     np.zeros(3, dtype=np.int64), dist, T,
                                     np.zeros(1, dtype=np.int64))
        out['ongrid_trace1_volumes'], out['ongrid_trace1_known'] = v1.copy(), k1.copy()
        out['ongrid_trace1_changed'] = np.int64(ch)
        # --- edge_check after that iteration -----------------------------------
        k2 = k1.copy()
        chk, e2 = refinement.edge_check(k2, rho, v1)
        out['ongrid_check_known'] = k2.copy()
        out['ongrid_check_counts'] = np.array([chk, e2], dtype=np.int64)
        # --- ongrid + full refine drivers --------------------------------------
        for tag, mode in (('changed3', ('changed', 3)), ('all_inf', ('all', -1)),
                          ('all2', ('all', 2))):
            v = vo.copy()
            th.refine('neargrid', mode, rho, v, dist, T, 1)
            out[f'ongrid_refine_{tag}'] = v
        # --- neargrid raw, then the default refine ------------------------------
        mxn, vn = th.bader_calc('neargrid', rho, fresh(), dist, T, 1)
        out['neargrid_maxima'], out['neargrid_raw_volumes'] = mxn, vn.copy()
        v = vn.copy()
        th.refine('neargrid', ('changed', 2), rho, v, dist, T, 1)
        out['neargrid_refine_changed2'] = v.copy()
        v = vn.copy()
        th.refine('neargrid', ('all', -1), rho, v, dist, T, 1)
        out['neargrid_refine_all_inf'] = v.copy()
        # --- sums, atoms, surface distance on the default result ---------------
        final = out['neargrid_refine_changed2']
        n = mxn.shape[0]
        q, vol = np.zeros(n), np.zeros(n)
        utils.charge_sum(q, vol, dV, rho, final)
        out['bader_charge'], out['bader_volume'] = q, vol
        if spin is not None:
            s, vol2 = np.zeros(n), np.zeros(n)
            utils.charge_sum(s, vol2, dV, spin, final)
            out['bader_spin'] = s
        frac = geo.maxima_fractional(mxn, shape, voxel_offset)
        cart = np.dot(frac, lattice)
        out['bader_maxima_cart'] = cart
        ba, bd, av = th.assign_to_atoms(cart, out['atoms'], out['lattice'], final, 1)
        out['bader_atoms'], out['bader_distance'], out['atoms_volumes'] = ba, bd, av
        na = out['atoms'].shape[0]
        q, vol = np.zeros(na), np.zeros(na)
        utils.charge_sum(q, vol, dV, rho, av)
        out['atoms_charge'], out['atoms_volume'] = q, vol
        vo_frac = np.dot(np.asarray(voxel_offset, float), geo.voxel_lattice(lattice, shape))
        sd = th.surface_distance(rho, av, out['lattice'], out['atoms'] - vo_frac, 1)
        out['atoms_surface_distance'] = (np.zeros(na) if sd is None else sd)
    return out


def cases():
    from pybader_b200 import synth
    # G1: three atoms, cubic, no vacuum
    c = synth.case_c1(24)
    rho, atoms = synth.make(c)
    yield 'g1_cubic24', dict(rho=rho, lattice=c['lattice'], atoms_cart=atoms)
    # G2: triclinic + vacuum + spin, anisotropic grid
    c = synth.case_triclinic((20, 24, 28), n_atoms=5, seed=7)
    c['sigmas'] = c['sigmas'] * 3.0
    rho, spin, atoms = synth.make(c, spin=True)
    yield 'g2_triclinic_vac', dict(rho=rho, lattice=c['lattice'], atoms_cart=atoms,
                                   vacuum_tol=1e-3, spin=spin,
                                   voxel_offset=(.5, .5, .5))
    # G3: quantised rocksalt -> exact ties between neighbours (tie-break order)
    c = synth.case_rocksalt(16, cells=2, offset=0.0, a=5.64)
    rho, atoms = synth.make(c)
    rho = np.round(rho, 1) + 0.05
    yield 'g3_ties16', dict(rho=rho, lattice=c['lattice'], atoms_cart=atoms)
    # G4: orthorhombic, unequal grid, vacuum with a large tolerance
    c = synth.case_slab((18, 20, 32), n_atoms=4, seed=11)
    c['sigmas'] = c['sigmas'] * 2.0
    rho, atoms = synth.make(c)
    yield 'g4_slab_vac', dict(rho=rho, lattice=c['lattice'], atoms_cart=atoms,
                              vacuum_tol=5e-2)


def main():
    ref = import_reference()
    for name, kw in cases():
        out = run_reference(ref, **kw)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'maxima ongrid/neargrid:', out['ongrid_maxima'].shape[0],
              out['neargrid_maxima'].shape[0], 'edges', int(out['ongrid_edges']),
              'trace1 changed', int(out['ongrid_trace1_changed']),
              'check', out['ongrid_check_counts'],
              '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
