"""pybader_b200 -- B200-native engine for the pybader hot path.

Host code is Python; all compute is hand-written CUDA for sm_100a behind the C
ABI of include/bader_b200.h (libbader_b200.so, loaded with ctypes).  There is
no CPU fallback.

    import pybader_b200
    pybader_b200.install()        # rebinds the names pybader.interface imported
    from pybader.interface import Bader   # unchanged reference front-end
"""
from . import geometry, synth  # noqa: F401

__version__ = "0.1.0"

_PATCHED = ('assign_to_atoms', 'bader_calc', 'dtype_calc', 'refine', 'surface_distance',
            'atom_assign', 'charge_sum', 'vacuum_assign', 'volume_mask')


def install(interface_module=None, readers=True):
    """Swap the numba hot path under the reference's `Bader` object for this
    engine by rebinding the names `pybader.interface` imported at
    interface.py:16-18.  With `readers` the CHGCAR / cube `read()` functions that
    `Bader.from_file` dispatches to (interface.py:141-173: `io_.read(...)` on the
    modules of `pybader.io`) are replaced too (pybader_b200/io).
    Returns the dict of replaced callables (for uninstall)."""
    from . import _lib, thread_handlers, utils
    _lib.load()   # fail loudly here, not in the middle of a run
    if interface_module is None:
        import pybader.interface as interface_module
    mine = dict(assign_to_atoms=thread_handlers.assign_to_atoms,
                bader_calc=thread_handlers.bader_calc, refine=thread_handlers.refine,
                surface_distance=thread_handlers.surface_distance, dtype_calc=utils.dtype_calc,
                atom_assign=utils.atom_assign, charge_sum=utils.charge_sum,
                vacuum_assign=utils.vacuum_assign, volume_mask=utils.volume_mask)
    old = {k: getattr(interface_module, k) for k in _PATCHED}
    for k, v in mine.items():
        setattr(interface_module, k, v)
    if readers:
        try:
            import pybader.io as ref_io
            from . import io as my_io
            old['io.vasp.read'], old['io.cube.read'] = ref_io.vasp.read, ref_io.cube.read
            ref_io.vasp.read, ref_io.cube.read = my_io.vasp.read, my_io.cube.read
        except ImportError:
            pass
    return old


def uninstall(old, interface_module=None):
    if interface_module is None:
        import pybader.interface as interface_module
    for k, v in old.items():
        if k.startswith('io.'):
            import pybader.io as ref_io
            setattr(getattr(ref_io, k.split('.')[1]), 'read', v)
        else:
            setattr(interface_module, k, v)
