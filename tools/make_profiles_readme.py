#!/usr/bin/env python
"""profiles/README.md from the committed bench lines, traffic capture and ncu summaries.

    python tools/make_profiles_readme.py r1v13 > profiles/README.md
"""
import glob
import json
import os
import sys

tag = sys.argv[1]
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles')


def load(name):
    with open(os.path.join(P, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


b = load(f'{tag}_bench_1024.json')
ref = load(f'{tag}_bench_ref.json')
with open(os.path.join(P, f'{tag}_traffic.json')) as f:
    tr = json.load(f)
N = 1024 ** 3
fam_k = {'stencil': ['k_seed_pointers'], 'resolve': ['k_tile_hist', 'k_tile_scan', 'k_tile_scatter', 'k_resolve_tiles'],
         'relabel': ['k_relabel_slots'], 'edge_eq': ['k_label_eq_bits'],
         'edge_flag': ['k_edge_from_eq', 'k_edge_deferred', 'k_eq_update'],
         'edge_dilate': ['k_edge_known'], 'trace': ['k_trace'], 'first': ['k_first_voxel_slots'],
         'edge_confirm': ['k_edge_confirm', 'k_edge_confirm_roots', 'k_edge_fix_clear', 'k_edge_fix_known',
                          'k_edge_fix_tomb', 'k_mark_interior'],
         'edge_check': ['k_filter_cached', 'k_inc_collect', 'k_inc_classify', 'k_inc_dilate', 'k_inc_mark',
                        'k_compact_known', 'k_bits_from_list', 'k_ec_init', 'k_ec_round', 'k_ec_collect_centres',
                        'k_ec_classify', 'k_ec_dilate', 'k_ec_finish']}
out = []
w = out.append
w(f"# profiles/ — measured evidence ({tag})\n")
w("Everything here was produced on one NVIDIA B200 (sm_100a, 148 SMs) through `gpurun` by "
  "`tools/gpu_round.sh` (single GPU) and the torchrun lines in `tools/gpu_scale.sh` (2/4/8 GPUs).  "
  "Numbers printed under ncu are never bench values: bench values come from `bench.py` (CUDA events on "
  "the library's stream), ncu supplies launch lists, DRAM traffic and the `--set full` details.\n")
w("## Headline (N = 1, 1024³, neargrid + refine('changed', 2))\n")
w("| quantity | value |\n|---|---|")
w(f"| step, density resident (`value`) | {b['ms_per_step']:.2f} ms → {b['value'] / 1e9:.2f} Gvoxel/s |")
e = b['e2e']
w(f"| end to end through `bdr_run`, host buffers (`e2e`) | {e['ms_per_step']:.1f} ms → {e['value'] / 1e9:.2f} Gvoxel/s "
  f"({e['h2d_bytes_per_step'] / 1e9:.2f} GB H2D + {e['d2h_bytes_per_step'] / 1e9:.2f} GB D2H per step; PCIe-bound) |")
h = e.get('handlers')
if h:
    w(f"| end to end through `thread_handlers.bader_calc` + `refine` (what the unmodified `Bader` drives; pageable numpy, "
      f"fresh session per step) | {h['ms_per_step']:.0f} ms → {h['value'] / 1e9:.2f} Gvoxel/s |")
c = b['cpu_baseline']
if c:
    w(f"| `cpu_baseline` (kind `{c['kind']}`, {c['cores']} host threads) | {c['value'] / 1e6:.2f} Mvoxel/s — {c['sample']} |")
rc = ref['cpu_baseline']
w(f"| reference arm (`--impl reference`, kind `{rc['kind']}`, {rc['cores']} host threads) | {ref['value'] / 1e6:.2f} Mvoxel/s — "
  f"{rc['sample']} |")
for tagw, label in (('c3', 'BASELINE config 3'), ('c4', 'BASELINE config 4')):
    fw = os.path.join(P, f'{tag}_bench_{tagw}.json')
    if os.path.exists(fw):
        dw = load(f'{tag}_bench_{tagw}.json')
        w(f"| {label} (`bench.py --workload {tagw}`) | {dw['ms_per_step']:.2f} ms → {dw['value'] / 1e9:.2f} Gvoxel/s; e2e "
          f"{dw['e2e']['ms_per_step']:.1f} ms; {dw['config']['maxima']} maxima, refine history {dw['refine_history_last_step']} |")
w(f"| kernels launched per step (`gpu_launches` / steps) | {b['gpu_launches'] / b['steps']:.0f}; host counter "
  f"read-backs per step: {b.get('host_syncs_per_step', float('nan')):.0f} |")
w(f"| SM clock during the timed region | {b['clocks']['sm_mhz']} MHz of {b['clocks']['sm_max_mhz']} (reasons: {b['clocks']['reasons'] or 'none'}) |")
r = b['roofline']
w(f"| `roofline` (dominant family: {r['kernel']}) | achieved {r['achieved']:.0f} GB/s algorithmic of {r['peak']:.0f} GB/s "
  f"({r['peak_source']}) = {r['frac']:.3f}; DRAM traffic per launch {((r.get('traffic') or 0) / 1e9):.2f} GB vs "
  f"{r.get('alg_bytes_per_launch', 0) / 1e9:.2f} GB algorithmic (gathers are served by L1/L2) |\n")
w("## Per kernel family, one step (CUDA events in bench.py; DRAM bytes from the ncu pass over every launch)\n")
w("| family | kernels | ms / step | launches | alg. B/voxel | achieved GB/s | frac of HBM peak | DRAM GB read+written (ncu) | alg. GB |")
w("|---|---|---|---|---|---|---|---|---|")
for name, k in sorted(b['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    ks = fam_k.get(name, [])
    dr = sum(tr['kernels'][x]['dram_read_bytes'] + tr['kernels'][x]['dram_write_bytes'] for x in ks if x in tr['kernels'])
    alg = k.get('alg_bytes_per_voxel')
    algb = (alg * N * k.get('passes_per_step', 1) / 1e9) if alg else (k.get('alg_bytes_per_launch', 0) * k['launches_per_step'] / 1e9)
    w(f"| {name} | {', '.join('`%s`' % x for x in ks if x in tr['kernels']) or '—'} | {k['ms_per_step']:.2f} | "
      f"{k['launches_per_step']:.0f} | {alg if alg else '—'} | "
      f"{k.get('achieved_gbs', 0):.0f} | {k.get('frac', 0):.3f} | {dr / 1e9:.2f} | {algb:.2f} |")
# share of the step per family: CUDA events (bench.py) vs the ncu launch list of the same command
import csv as _csv
rows = list(_csv.reader(open(os.path.join(P, f'{tag}_launches_1024.csv'))))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
launches = []
for r in rows[h + 1:]:
    if len(r) > mv:
        name = r[kn].replace('void ', '').split('(')[0].split('<')[0]
        t = float(r[mv].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[mu][:2].rstrip('e'), 1e-6)
        launches.append((name, t))
first = [i for i, (n_, _) in enumerate(launches) if n_ in ('k_seed_pointers', 'k_ongrid_pointers')]
last = launches[first[-1]:]                      # the timed step (the last one of the run)
k2f = {k: f for f, ks in fam_k.items() for k in ks}
ncu_ms = {}
for n_, t in last:
    f = k2f.get(n_)
    if f:
        ncu_ms[f] = ncu_ms.get(f, 0.0) + t
tot_ncu, tot_ev = sum(ncu_ms.values()), sum(k['ms_per_step'] for k in b['kernels'].values())
w("\nShare of the step, CUDA events in bench.py vs the ncu launch list of the same command "
  f"(`{tag}_launches_1024.csv`, last step; ncu serialises launches and runs them cold, so only shares compare):\n")
w("| family | events ms | events share | ncu ms | ncu share |\n|---|---|---|---|---|")
for name, k in sorted(b['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']):
    w(f"| {name} | {k['ms_per_step']:.2f} | {k['ms_per_step'] / tot_ev:.3f} | {ncu_ms.get(name, 0):.2f} | "
      f"{ncu_ms.get(name, 0) / tot_ncu:.3f} |")
w(f"\nSum of the families: {sum(k['ms_per_step'] for k in b['kernels'].values()):.2f} ms of the {b['ms_per_step']:.2f} ms step; "
  "the rest is host round trips between data-dependent launches (counter read-backs).\n")
w("## Scaling (weak: 2^30 voxels per GPU; x-slabs, NCCL + NVLink peer loads)\n")
w("| GPUs | grid | ms / step | Gvoxel/s | e2e Gvoxel/s | kernels on rank 0, ms | note |")
w("|---|---|---|---|---|---|---|")
w(f"| 1 | 1024³ | {b['ms_per_step']:.2f} | {b['value'] / 1e9:.1f} | {b['e2e']['value'] / 1e9:.2f} | "
  f"{sum(k['ms_per_step'] for k in b['kernels'].values()):.1f} | refine ('changed', 2) |")
for n in (2, 4, 8):
    f = os.path.join(P, f'{tag}_bench_n{n}.json')
    if not os.path.exists(f):
        continue
    d = load(f'{tag}_bench_n{n}.json')
    km = d.get('kernel_ms_per_step', {})
    w(f"| {n} | {d['config']['workload'].split(' ')[0]} | {d['ms_per_step']:.2f} | {d['value'] / 1e9:.1f} | "
      f"{d['e2e']['value'] / 1e9:.2f} | {sum(k['ms_per_step'] for k in d['kernels'].values()):.1f} "
      f"(slowest rank {km.get('max_over_ranks', 0):.1f}) | "
      f"refine {tuple(d['config']['refine_mode'])}, history {d['refine_history_last_step']}; {d.get('neargrid_passes')} rounds, "
      f"{d.get('exit_rounds')} exit rounds; efficiency vs N=1 {b['ms_per_step'] / d['ms_per_step']:.2f} |")
w("\nEvery N runs the same algorithm (`refine(('changed', 2))`); the round loops run inside the library over its own NCCL "
  "communicator (DESIGN.md section 7).  The gap between the kernel sum and the step is exit resolution / numbering in "
  "torch, the wait for the slowest rank of every round and the host decisions per round.  Parity of the sharded runs: "
  "`*_sharded_w{2,4,8}_pytest.log`, `*_sharded_handlers_w2_pytest.log`.\n")
w("## Files\n")
for f in sorted(os.listdir(P)):
    if f == 'README.md':
        continue
    what = {'bench_1024.json': "bench.py line, N=1", 'bench_ref.json': "bench.py --impl reference line",
            'launches_1024.csv': "ncu launch list (gpu__time_duration.sum) of `bench.py --steps 1 --warmup 3`",
            'traffic.json': "DRAM bytes and duration of every launch of one step (tools/ncu_traffic.py)",
            'ncu_set_full_1024.txt': "key metrics of the `--set full` captures of the top kernels (tools/ncu_summary.py)",
            'pytest_gpu.log': "`pytest -m gpu -s -rs` on a one-GPU box (the multi-GPU cases skip there)",
            'smoke.log': "`__graft_entry__.smoke()`",
            'bench_c3.json': "bench.py --workload c3 (BASELINE config 3 shape)",
            'bench_c4.json': "bench.py --workload c4 (BASELINE config 4 shape)",
            'sharded_w2_pytest.log': "`pytest tests/test_gpu_sharded.py` under `gpurun --gpus 2`: 2 ranks vs 1 GPU, library and Python loops, unmodified Bader over 2 ranks",
            'sharded_w4_pytest.log': "the same under `gpurun --gpus 4` (2- and 4-rank cases)",
            'sharded_w8_pytest.log': "the 8-rank case under `gpurun --gpus 8`",
            'sharded_handlers_w2_pytest.log': "first run of the unmodified `Bader.__call__` over 2 ranks",
            'trace_occupancy.txt': "A/B of k_trace launch bounds and refill chunk",
            'sanitizer_memcheck.log': "compute-sanitizer memcheck over smoke(): 0 errors",
            'sanitizer_memcheck_parity.log': "compute-sanitizer memcheck over the vacuum / ragged parity cases: 0 errors",
            'sanitizer_racecheck.log': "compute-sanitizer racecheck over smoke(): the intended in-tile chase race only (DESIGN.md section 8)",
            'io_bench_256.json': "tools/io_bench.py: CHGCAR reader (GPU text -> grid) next to the reference's conversion"}
    desc = next((v for k, v in what.items() if f.endswith(k)), '')
    if '_bench_n' in f:
        desc = "bench.py line under torchrun, N = " + f.split('_bench_n')[1].split('.')[0]
    w(f"* `{f}` — {desc or 'earlier capture of this round, kept for the history of the kernels'}")
print('\n'.join(out))
