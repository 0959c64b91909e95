#!/usr/bin/env python
"""bench.py -- neargrid + refine voxels/s on synthetic Gaussian-superposition
densities (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one resident density:
    labels := 0 ; bader_calc('neargrid') ; refine('neargrid', ('changed', 2))
i.e. thread_handlers.bader_calc + thread_handlers.refine of the reference
(thread_handlers.py:15-75, 128-236), the path BASELINE.json's metric names.

`value`  = voxels / device time of the step, density already resident in HBM.
`e2e`    = the same metric through the C ABI one-shot call `bdr_run` with HOST
           buffers: pinned-host density in, narrowed labels + maxima out, the
           H2D / D2H copies inside the timed region.  `e2e.handlers` is the same
           step through the reference-shaped Python handlers (what the unmodified
           `Bader` object drives) with pageable numpy arrays.
`roofline` is for the family of kernels that takes the largest share of the
step, from CUDA events recorded on the library's own stream around every launch.
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference (pybader's
numba thread handlers from baseline/_ref, all host threads) on a bounded sample
of the same workload family; only if pybader cannot be imported on the box the
C port of its algorithm (oracle/) runs instead, and the line says so.
`--workload c3|c4` measures BASELINE configs 3 and 4 instead of the 1024^3 target.
N > 1 (under torchrun): pybader_b200.sharded.bench, weak scaling, 2^30 voxels per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "neargrid+refine voxels/s"
UNIT = "voxels/s"
VOXEL = 22.5 / 1024.0      # Angstrom per voxel; BASELINE config 5 is 45 A over 2048
SPACING = 204.8            # voxels between neighbouring atoms (1000 atoms in 2048^3)

# weak scaling: 2^30 voxels per GPU; N=1 is the 1024^3 north-star target, N=8
# is BASELINE config 5 (2048^3, 1000 atoms)
SHAPES = {1: (1024, 1024, 1024), 2: (2048, 1024, 1024), 4: (2048, 2048, 1024),
          8: (2048, 2048, 2048)}

# algorithmic bytes per voxel of each streaming kernel family
# edge_eq: labels R 4 + four bit volumes W 4/8; edge_flag (bits -> candidate bits): R 4/8 + W 1/8
ALG_BYTES = {'stencil': 12, 'resolve': 8, 'relabel': 8, 'edge_eq': 4.5, 'edge_flag': 0.625,
             'edge_dilate': 1.25, 'first': 4, 'charge_sum': 12, 'vacuum': 12, 'narrow': 5}


def workload_case(shape):
    from pybader_b200 import synth
    cells = tuple(max(1, int(round(s / SPACING))) for s in shape)
    a = tuple(s * VOXEL for s in shape)
    return synth.case_lattice_sites(shape, cells, a, seed=2048), cells


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.active = True      # rows are kept only while a timed region runs

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            if self.active:
                self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------- CPU legs ----
def cpu_sample_inputs(n):
    """bounded sample of the workload family: n^3 periodic cell, same voxel size,
    2^3 jittered atoms"""
    from pybader_b200 import geometry as geo, synth
    c = synth.case_lattice_sites((n, n, n), (2, 2, 2), (n * VOXEL,) * 3, seed=2048)
    tx, ty, tz = synth.separable_tables(c)
    rho = np.einsum('ai,aj,ak->ijk', tx, ty, tz, optimize=True)
    rho = np.ascontiguousarray(rho)
    return rho, geo.distance_matrix(c['lattice'], rho.shape), geo.T_grad(c['lattice'], rho.shape)


def cpu_step(orc, rho, dist, T):
    vol = np.zeros(rho.shape, dtype=np.int32)
    _, vol = orc.bader_calc('neargrid', rho, vol, dist, T)
    orc.refine('neargrid', ('changed', 2), rho, vol, dist, T)
    return vol


def cpu_baseline_leg(n=256, port=False):
    """the reference beside the GPU number: the unmodified pybader (numba, all host
    threads) on a bounded sample, else the single-threaded C port"""
    if not port:
        ref, why = load_numba_reference()
        if ref is not None:
            cores = os.cpu_count() or 1
            sp = 96.0
            rho, dist, T = ref_sample_inputs(sp)
            t0 = time.perf_counter()
            numba_step(ref, rho, dist, T, cores)
            dt = time.perf_counter() - t0
            return {"value": rho.size / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"{rho.shape[0]}^3 periodic cell, 2x2x2 jittered atoms of the same workload "
                              f"family ({sp:.0f}-voxel atom spacing), unmodified pybader thread handlers "
                              f"bader_calc('neargrid') + refine(('changed', 2)), threads={cores}, JIT warm, "
                              f"{dt:.1f} s"}
    from oracle import pyoracle as orc
    rho, dist, T = cpu_sample_inputs(n)
    cpu_step(orc, rho[:32, :32, :32].copy(), dist, T)          # load / warm the .so
    t0 = time.perf_counter()
    cpu_step(orc, rho, dist, T)
    dt = time.perf_counter() - t0
    return {"value": rho.size / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n}^3 periodic cell of the same workload family (voxel {VOXEL:.5f} A, "
                      f"8 jittered atoms), oracle/bader_oracle.c neargrid + refine('changed',2), "
                      f"{dt:.1f} s on one host core"}


def ref_sample_inputs(spacing, n_atoms_axis=2):
    """bounded sample for the reference arm: a periodic cubic cell of 2x2x2 jittered atoms of
    the bench workload family (same voxel size, same Gaussian widths relative to the atom
    spacing, every atom with six neighbours like in the 5x5x5 / 10x10x10 cells of the GPU
    arm), `spacing` voxels between atoms"""
    from pybader_b200 import geometry as geo, synth
    n = int(round(n_atoms_axis * spacing))
    c = synth.case_lattice_sites((n, n, n), (n_atoms_axis,) * 3, (n * VOXEL,) * 3, seed=2048)
    tx, ty, tz = synth.separable_tables(c)
    rho = np.ascontiguousarray(np.einsum('ai,aj,ak->ijk', tx, ty, tz, optimize=True))
    return rho, geo.distance_matrix(c['lattice'], rho.shape), geo.T_grad(c['lattice'], rho.shape)


def numba_step(ref, rho, dist, T, threads):
    """the reference's own hot path, unmodified: thread_handlers.bader_calc('neargrid') +
    thread_handlers.refine('neargrid', ('changed', 2)) (thread_handlers.py:15-75, 128-236)"""
    from baseline.refload import quiet
    vol = np.zeros(rho.shape, dtype=ref['utils'].dtype_calc(-rho.size))
    with quiet():
        mx, vol = ref['th'].bader_calc('neargrid', rho, vol, dist, T, threads)
        ref['th'].refine('neargrid', ('changed', 2), rho, vol, dist, T, threads)
    return vol


def load_numba_reference():
    """(ref modules, None) or (None, why not)"""
    try:
        from baseline.refload import import_reference
        ref = import_reference()
        # JIT warm-up of every signature the step uses, on a 16^3 cell
        rho, dist, T = ref_sample_inputs(8)
        numba_step(ref, rho, dist, T, 1)
        numba_step(ref, rho, dist, T, 2)
        return ref, None
    except Exception as e:                      # numba / pybader missing on this box
        return None, f"{type(e).__name__}: {e}"


def pick_spacing(rate, steps, budget_s):
    """largest atom spacing (voxels) <= the GPU arm's whose 2x2x2-atom cell lets `steps`
    steps fit the time budget at `rate` voxels/s"""
    for sp in (SPACING, 176.0, 160.0, 144.0, 128.0, 112.0, 96.0, 80.0, 64.0):
        if steps * (2 * sp) ** 3 / rate <= budget_s:
            return sp
    return 48.0


def reference_arm(args):
    """--impl reference: the UNMODIFIED reference (pybader's numba thread handlers from
    baseline/_ref) on all host cores, on a bounded sample of the GPU arm's workload family.
    If pybader / numba cannot be imported on this box the C port of the same algorithm
    (oracle/) runs instead, one independent cell per host thread, and the line says so."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    shape = SHAPES.get(args.gpus, SHAPES[1])
    ref, why = (None, 'disabled by --ref-port') if args.ref_port else load_numba_reference()
    total_steps = args.steps + args.warmup
    if ref is not None:
        kind = 'reference'
        rho, dist, T = ref_sample_inputs(48)
        t0 = time.perf_counter()
        numba_step(ref, rho, dist, T, cores)
        rate = rho.size / (time.perf_counter() - t0)
        sp = args.ref_spacing or pick_spacing(rate * 1.5, total_steps, args.ref_budget)
        rho, dist, T = ref_sample_inputs(sp)
        if not args.ref_spacing:
            # the small cell under-estimates the rate of a large one: time one step of the chosen
            # cell and move one notch up or down (at most three times) to fill the budget
            ladder = [SPACING, 176.0, 160.0, 144.0, 128.0, 112.0, 96.0, 80.0, 64.0, 48.0]
            for _ in range(3):
                t0 = time.perf_counter()
                numba_step(ref, rho, dist, T, cores)
                t1 = time.perf_counter() - t0
                i = ladder.index(sp)
                if t1 * total_steps > 1.3 * args.ref_budget and i + 1 < len(ladder):
                    sp = ladder[i + 1]
                elif i > 0 and t1 * total_steps * (ladder[i - 1] / sp) ** 3 < 0.9 * args.ref_budget:
                    sp = ladder[i - 1]
                else:
                    break
                rho, dist, T = ref_sample_inputs(sp)
        n_vox = rho.size

        def one_step():
            numba_step(ref, rho, dist, T, cores)

        sample = (f"one {rho.shape[0]}^3 periodic cell per step: 2x2x2 jittered atoms of the same "
                  f"workload family ({sp:.1f}-voxel atom spacing; the GPU arm's is {SPACING}), "
                  f"unmodified pybader {ref['root']} thread_handlers.bader_calc('neargrid') + "
                  f"refine('neargrid', ('changed', 2)), threads={cores}, numba JIT warmed on 16^3")
    else:
        kind = 'port'
        from oracle import pyoracle as orc
        n = args.ref_sample
        rho, dist, T = cpu_sample_inputs(n)
        bricks = [rho.copy() for _ in range(cores)]
        cpu_step(orc, rho[:32, :32, :32].copy(), dist, T)
        n_vox = cores * rho.size

        def one_step():
            ths = [threading.Thread(target=cpu_step, args=(orc, b, dist, T)) for b in bricks]
            for t in ths:
                t.start()
            for t in ths:
                t.join()

        sample = (f"pybader not importable here ({why}); C port instead: {cores} independent {n}^3 "
                  f"cells per step (one host thread each), same workload family; oracle/ port of "
                  f"methods.neargrid + thread_handlers.refine('changed',2)")
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n_vox / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(shape), "method": "neargrid",
                   "refine_method": "neargrid", "refine_mode": ["changed", 2]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(shape, mode="('changed',2)"):
    return (f"{shape[0]}x{shape[1]}x{shape[2]} periodic orthorhombic cell, jittered simple-lattice "
            f"Gaussian superposition ({SPACING:.1f}-voxel atom spacing), neargrid + "
            f"refine{mode}")


# kernel names behind each family (for the DRAM traffic measured by ncu, profiles/*_traffic.json)
FAMILY_KERNELS = {'trace': ['k_trace'], 'stencil': ['k_seed_pointers', 'k_ongrid_pointers'],
                  'resolve': ['k_resolve_tiles', 'k_tile_hist', 'k_tile_scan', 'k_tile_scatter'],
                  'edge_eq': ['k_label_eq_bits'], 'edge_flag': ['k_edge_from_eq', 'k_edge_deferred', 'k_eq_update'],
                  'edge_dilate': ['k_edge_known'], 'relabel': ['k_relabel_slots'],
                  'first': ['k_first_voxel_slots'], 'edge_confirm': ['k_edge_confirm']}


# kernels launched per pass over the grid in the families that chain several kernels
LAUNCHES_PER_PASS = {'resolve': 4}


def measured_traffic(family, n_voxels):
    """DRAM bytes per launch of a kernel family from the committed ncu capture of one
    step of this workload (dram__bytes_read.sum + dram__bytes_write.sum), or None"""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json')), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
        except (OSError, ValueError):
            continue
        if d.get('voxels') != n_voxels:
            continue
        tot, n = 0.0, 0
        for k in FAMILY_KERNELS.get(family, []):
            e = d['kernels'].get(k)
            if e:
                tot += e['dram_read_bytes'] + e['dram_write_bytes']
                n += e['launches']
        if n:
            n /= LAUNCHES_PER_PASS.get(family, 1)
            return {"bytes_per_launch": tot / n, "launches_captured": n,
                    "source": 'profiles/' + os.path.basename(path)}
    return None


def kernel_accounting(prof, n_voxels, steps, tsteps, tvox, ms_per_step):
    """per-family table (CUDA-event times on the library's stream) and the roofline
    object of the family that takes the largest share of the step; n_voxels is what
    one launch of a streaming kernel covers (the rank's grid or window)"""
    peak, peak_src = peaks()
    kernels = {}
    for name, (ms, n) in prof.items():
        k = {"ms_per_step": ms / steps, "launches_per_step": n / steps}
        if name in ALG_BYTES:
            gb = ALG_BYTES[name] * n_voxels * 1e-9
            passes = n / LAUNCHES_PER_PASS.get(name, 1)   # full passes over the grid
            if name == 'edge_flag' and 'edge_dilate' in prof:
                passes = prof['edge_dilate'][1]           # 2-3 small launches per pass; k_edge_known runs once
            k["alg_bytes_per_voxel"] = ALG_BYTES[name]
            k["passes_per_step"] = passes / steps
            k["achieved_gbs"] = gb / (ms / passes * 1e-3)
            k["frac"] = k["achieved_gbs"] / peak
        kernels[name] = k
    if 'trace' in kernels and tvox:
        # gather-bound: 7 fp64 gathers + 1 known byte per trajectory step, label R+W per voxel
        tb = tsteps * (7 * 8 + 1) + tvox * (4 + 4 + 4 + 1)
        kernels['trace'].update({"alg_bytes_per_launch": tb / prof['trace'][1],
                                 "achieved_gbs": tb * 1e-9 / (prof['trace'][0] * 1e-3),
                                 "steps_per_voxel": tsteps / tvox,
                                 "voxels_per_step": tvox / steps})
        kernels['trace']["frac"] = kernels['trace']["achieved_gbs"] / peak
    dom = max((k for k in kernels if 'achieved_gbs' in kernels[k]),
              key=lambda k: kernels[k]['ms_per_step'])
    dk = kernels[dom]
    tr = measured_traffic(dom, n_voxels)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": dk['achieved_gbs'], "peak": peak,
                "unit": "GB/s", "frac": dk['achieved_gbs'] / peak,
                "traffic": tr and tr['bytes_per_launch'], "traffic_source": tr and tr['source'],
                "alg_bytes_per_launch": dk['achieved_gbs'] * 1e9 * 1e-3 * dk['ms_per_step']
                / dk.get('passes_per_step', dk['launches_per_step']),
                "peak_source": peak_src,
                "ms_per_launch": dk['ms_per_step'] / dk.get('passes_per_step', dk['launches_per_step']),
                "share_of_step": dk['ms_per_step'] / ms_per_step}
    return kernels, roofline


# ---------------------------------------------------------------- GPU arm ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=0, help='override: cubic N^3 single-GPU grid')
    ap.add_argument('--workload', default='target', choices=['target', 'c3', 'c4'],
                    help='target: the 1024^3 north-star cell (default); c3 / c4: BASELINE configs 3 / 4 '
                         '(extra lines kept under profiles/)')
    ap.add_argument('--cpu-sample', type=int, default=256)
    ap.add_argument('--ref-sample', type=int, default=160, help='port fallback: cell edge')
    ap.add_argument('--ref-spacing', type=float, default=0.0,
                    help='reference arm: atom spacing in voxels of its 2x2x2-atom cell (0: picked to fit --ref-budget)')
    ap.add_argument('--ref-budget', type=float, default=120.0, help='reference arm: seconds for all steps')
    ap.add_argument('--ref-port', action='store_true', help='reference arm / cpu baseline: force the C port')
    ap.add_argument('--halo', type=int, default=4)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        return reference_arm(args)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun")
    if world > 1:
        from pybader_b200 import sharded
        return sharded.bench(args, rank, world, local)

    from pybader_b200 import build, geometry as geo, synth
    build.build()
    from pybader_b200.engine import Engine, LABELS_BADER

    vac_tol = None
    if args.workload == 'c3':
        # BASELINE config 3: triclinic 128-atom cell 360x360x480, vacuum_tol 1e-3
        case = synth.case_triclinic((360, 360, 480), n_atoms=128, seed=1234)
        shape, vac_tol = case['shape'], 1e-3
        name = "BASELINE config 3: 360x360x480 triclinic 128-atom cell, vacuum_tol 1e-3, neargrid + refine('changed',2)"
    elif args.workload == 'c4':
        # BASELINE config 4: 512x512x1024 orthorhombic slab with vacuum (the spin density rides along
        # in the sums only, which are outside the metric's path)
        case = synth.case_slab((512, 512, 1024), n_atoms=64, seed=4321)
        shape, vac_tol = case['shape'], 1e-3
        name = "BASELINE config 4: 512x512x1024 slab with vacuum, vacuum_tol 1e-3, neargrid + refine('changed',2)"
    else:
        shape = (args.size,) * 3 if args.size else SHAPES[1]
        case, cells = workload_case(shape)
        name = workload_name(shape)
    N = int(np.prod(shape))
    dist = geo.distance_matrix(case['lattice'], shape)
    T = geo.T_grad(case['lattice'], shape)
    dV = geo.voxel_volume(case['lattice'], shape)
    e = Engine(shape, device=local)
    if args.workload == 'c3':
        e.synth_general(0, case['lattice'], case['frac_atoms'], case['amps'], case['sigmas'])
    else:
        e.synth_separable(0, *synth.separable_tables(case))
    n_atoms = len(case['amps'])
    mode = ('changed', 2)

    def step():
        e.clear_labels(LABELS_BADER)
        if vac_tol is not None:
            e.vacuum_assign(vac_tol, dV)       # Bader.volumes_init (interface.py:449-469)
        mx = e.bader_calc('neargrid', dist, T)
        hist = e.refine(LABELS_BADER, mode[0], mode[1], dist, T)
        return mx, hist

    for _ in range(args.warmup):
        mx, hist = step()
    n_max = mx.shape[0]

    # ---- timed region: K steps, CUDA events on the library's stream ----------
    clocks = ClockSampler(local)
    clocks.start()
    e.profile(True)
    e.profile_reset()
    l0, s0 = e.launch_count(), e.sync_count()
    e.synchronize()
    e.timer_start()
    for _ in range(args.steps):
        step()
    total_ms = e.timer_stop()
    e.synchronize()
    launches = e.launch_count() - l0
    host_syncs = e.sync_count() - s0
    prof = e.profile_get()
    tsteps, tvox = e.trace_steps()
    e.profile(False)
    clocks.active = False
    ms_per_step = total_ms / args.steps
    value = N / (ms_per_step * 1e-3)

    # ---- per-kernel accounting and the roofline object ---------------------
    kernels, roofline = kernel_accounting(prof, N, args.steps, tsteps, tvox, ms_per_step)

    # ---- e2e: host buffers through bdr_run -------------------------------------
    e2e = None
    if not args.no_e2e:
        import torch
        from pybader_b200.utils import dtype_calc
        ldt = np.dtype(dtype_calc(-n_max))
        host_rho = torch.empty(N, dtype=torch.float64, pin_memory=True).numpy().reshape(shape)
        host_lab = torch.empty(N * ldt.itemsize, dtype=torch.uint8,
                               pin_memory=True).numpy().view(ldt).reshape(shape)
        from pybader_b200._lib import check
        check(e.lib.bdr_download_density(e.h, 0, host_rho.ctypes.data))
        cap = max(1 << 12, 2 * n_max)
        e.run(host_rho, vac_tol, dV, 'neargrid', mode[0], mode[1], dist, T, ldt, cap,
              want_sums=False, out_labels=host_lab)
        e.synchronize()
        clocks.active = True
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e.run(host_rho, vac_tol, dV, 'neargrid', mode[0], mode[1], dist, T, ldt, cap,
                  want_sums=False, out_labels=host_lab)
        e.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        clocks.active = False
        e2e = {"value": N / dt, "unit": UNIT, "h2d_bytes_per_step": N * 8,
               "d2h_bytes_per_step": N * ldt.itemsize + n_max * 24, "ms_per_step": dt * 1e3,
               "api": "bdr_run (C ABI, pinned host density in, narrowed labels + maxima out)"}
        # second figure: the path the unmodified `Bader` object drives -- thread_handlers.bader_calc
        # + thread_handlers.refine with ordinary (pageable) numpy arrays, a fresh session per step
        # like a fresh Bader.__call__: content hash + upload of the density, fresh int32 volumes
        # (interface.py:456-457), narrowed labels back, refine's labels back if anything changed
        from pybader_b200 import session, thread_handlers as th, utils as ut
        rho_np = np.array(host_rho)          # pageable copy
        del host_rho, host_lab

        def handler_step():
            session.close_all()
            vol = np.zeros(shape, dtype=ut.dtype_calc(-N))
            if vac_tol is not None:
                vol, _, _ = ut.vacuum_assign(rho_np, vol, np.float64(vac_tol), rho_np, dV)
            mx_, vol = th.bader_calc('neargrid', rho_np, vol, dist, T, 1)
            th.refine('neargrid', mode, rho_np, vol, dist, T, 1)
            return vol

        e.close()
        handler_step()
        hsteps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(hsteps):
            vol_h = handler_step()
        dth = (time.perf_counter() - t0) / hsteps
        e2e["handlers"] = {"value": N / dth, "unit": UNIT, "ms_per_step": dth * 1e3, "steps": hsteps,
                           "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * vol_h.dtype.itemsize,
                           "api": "pybader_b200.thread_handlers.bader_calc + refine (the names "
                                  "pybader.interface binds), pageable numpy arrays, fresh session per step"}
        session.close_all()
        del rho_np, vol_h

    clock_info = clocks.stop()      # sampled through both timed regions (device steps and e2e)
    cpu = None if args.no_cpu else cpu_baseline_leg(args.cpu_sample, args.ref_port)
    e.close()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "atoms": n_atoms, "maxima": n_max,
                   "method": "neargrid", "refine_method": "neargrid", "refine_mode": list(mode),
                   "l2": "inputs (8 B/voxel density) are far larger than the 126 MB L2; no flush",
                   "parallelism": "1 GPU"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "host_syncs_per_step": host_syncs / args.steps,
        "clocks": clock_info, "kernels": kernels,
        "refine_history_last_step": hist,
    }
    print(json.dumps(out))


if __name__ == '__main__':
    main()
