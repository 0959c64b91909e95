"""bench.py contract on the CPU: the reference arm runs without a GPU and prints the
JSON line the driver expects; the accounting helpers are consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_reference_line(d, kind):
    assert d['impl'] == 'reference' and d['metric'] == 'neargrid+refine voxels/s'
    assert d['unit'] == 'voxels/s' and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['value'] > 0 and d['steps'] == 1 and d['vs_baseline'] is None
    cb = d['cpu_baseline']
    assert cb['kind'] == kind and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {"value": d['value'], "unit": 'voxels/s', "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert '1024x1024x1024' in d['config']['workload']


def test_reference_arm_line_port_fallback():
    """the C port (what runs when pybader / numba cannot be imported on the box)"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1', '--ref-sample', '40', '--ref-port'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_reference_line(d, 'port')
    assert 'not importable' in d['cpu_baseline']['sample']


def test_reference_arm_line_real_pybader():
    """the unmodified reference (numba thread handlers) when it is importable here"""
    sys.path.insert(0, ROOT)
    from baseline.refload import find_reference
    if find_reference() is None:
        import pytest
        pytest.skip("no pybader in baseline/_ref or /root/reference")
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '1', '--ref-spacing', '20'],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_reference_line(d, 'reference')
    assert 'unmodified pybader' in d['cpu_baseline']['sample'] and 'threads=' in d['cpu_baseline']['sample']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--gpus', '2', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_kernel_accounting_per_pass():
    sys.path.insert(0, ROOT)
    import bench as B
    N = 1024 ** 3
    prof = {'stencil': (15.4, 2), 'edge_eq': (2.8, 2), 'edge_flag': (1.2, 14), 'edge_dilate': (4.6, 6),
            'resolve': (4.0, 8), 'trace': (35.0, 26)}
    kernels, roof = B.kernel_accounting(prof, N, 2, 2 * 239_000_000, 2 * 79_000_000, 42.0)
    assert abs(kernels['stencil']['ms_per_step'] - 7.7) < 1e-9
    assert kernels['edge_flag']['passes_per_step'] == 3 and kernels['resolve']['passes_per_step'] == 1
    assert kernels['edge_eq']['passes_per_step'] == 1
    gbs = 4.5 * N * 1e-9 / (2.8 / 2 * 1e-3)
    assert abs(kernels['edge_eq']['achieved_gbs'] - gbs) < 1e-6 * gbs
    assert roof['kernel'] == 'trace' and 0 < roof['frac'] < 1 and roof['bound'] == 'hbm'
    assert abs(roof['share_of_step'] - 17.5 / 42.0) < 1e-9
