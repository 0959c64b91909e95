import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    g = os.path.join(ROOT, 'tests', 'golden')
    return sorted(f[:-4] for f in os.listdir(g) if f.endswith('.npz'))


@pytest.fixture(scope='session', params=golden_names())
def golden(request):
    import numpy as np
    path = os.path.join(ROOT, 'tests', 'golden', request.param + '.npz')
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    d['name'] = request.param
    tol = float(d['vacuum_tol'])
    d['vacuum_tol'] = None if tol != tol else tol
    return d
