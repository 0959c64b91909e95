"""Locate and import the UNMODIFIED reference (pybader, numba) for the reference arm
of bench.py, the golden generators and the boundary tests.

Search order: baseline/_ref (pip --target install of /root/reference, made in the build
container; travels to the GPU box), then /root/reference (build container only).
SURVEY.md section 8c recipe: `import pybader` runs a first-run hook unless a config.ini
exists, so HOME points at a scratch directory that holds the reference's own DEFAULT and
`speed` profiles (entry_points.py:326-345); numba's cache goes to a writable directory.
Nothing under pybader_b200/ imports this module.
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

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


def find_reference():
    for root in (os.path.join(HERE, '_ref'), '/root/reference'):
        if os.path.isfile(os.path.join(root, 'pybader', 'interface.py')):
            return root
    return None


def import_reference(scratch=None):
    """dict of the reference's modules + `Bader`, or raises ImportError with the reason."""
    root = find_reference()
    if root is None:
        raise ImportError("pybader not found in baseline/_ref or /root/reference")
    scratch = scratch or os.environ.get('PYBADER_REF_HOME', '/tmp/pybader_ref_home')
    cfg = os.path.join(scratch, '.config', 'bader')
    os.makedirs(cfg, exist_ok=True)
    ini = os.path.join(cfg, 'config.ini')
    if not os.path.exists(ini):
        with open(ini, 'w') as f:
            f.write(CONFIG_INI)
    os.environ['HOME'] = scratch
    os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(scratch, 'numba_cache'))
    if root not in sys.path:
        sys.path.insert(0, root)
    import pybader  # noqa: F401
    from pybader import interface, methods, refinement, thread_handlers, utils
    return dict(root=root, interface=interface, methods=methods, refinement=refinement,
                th=thread_handlers, utils=utils, Bader=interface.Bader)


@contextlib.contextmanager
def quiet():
    """tqdm bars and progress prints of the reference go to stdout (utils.py:140)"""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
