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
         'relabel': ['k_relabel_slots'], 'edge_flag': ['k_label_eq_bits', 'k_edge_from_eq', 'k_edge_deferred'],
         'edge_dilate': ['k_edge_known'], 'trace': ['k_trace'], 'first': ['k_first_voxel_slots'],
         'edge_confirm': ['k_edge_confirm', 'k_edge_fix_clear', 'k_edge_fix_known', 'k_edge_fix_tomb'],
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
c = b['cpu_baseline']
if c:
    w(f"| CPU oracle port, 1 core (`cpu_baseline`) | {c['value'] / 1e6:.2f} Mvoxel/s ({c['sample'].split(',')[0]}) |")
w(f"| reference arm (`--impl reference`, {ref['cpu_baseline']['cores']} host threads) | {ref['value'] / 1e6:.1f} Mvoxel/s |")
w(f"| kernels launched per step (`gpu_launches` / steps) | {b['gpu_launches'] / b['steps']:.0f} |")
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
    w(f"| {n} | {d['config']['workload'].split(' ')[0]} | {d['ms_per_step']:.2f} | {d['value'] / 1e9:.1f} | "
      f"{d['e2e']['value'] / 1e9:.2f} | {sum(k['ms_per_step'] for k in d['kernels'].values()):.1f} | "
      f"refine ('all', 2): one more full edge pass + trace than N = 1; {d.get('neargrid_passes')} rounds, "
      f"{d.get('exit_rounds')} exit rounds |")
w("\nThe N > 1 step is a heavier algorithm than the N = 1 step (DESIGN.md section 7) and is driven from Python "
  "with an all-reduce per round; the gap between the kernel sum and the step is that protocol.\n")
w("## Files\n")
for f in sorted(os.listdir(P)):
    if f == 'README.md':
        continue
    what = {'bench_1024.json': "bench.py line, N=1", 'bench_ref.json': "bench.py --impl reference line",
            'launches_1024.csv': "ncu launch list (gpu__time_duration.sum) of `bench.py --steps 1 --warmup 3`",
            'traffic.json': "DRAM bytes and duration of every launch of one step (tools/ncu_traffic.py)",
            'ncu_set_full_1024.txt': "key metrics of the `--set full` captures of the top kernels (tools/ncu_summary.py)",
            'pytest_gpu.log': "`pytest -m gpu` on the box", 'smoke.log': "`__graft_entry__.smoke()`",
            'sanitizer_memcheck.log': "compute-sanitizer memcheck over smoke(): 0 errors",
            'sanitizer_racecheck.log': "compute-sanitizer racecheck over smoke(): the intended in-tile chase race only (DESIGN.md section 8)",
            'io_bench_256.json': "tools/io_bench.py: CHGCAR reader (GPU text -> grid) next to the reference's conversion"}
    desc = next((v for k, v in what.items() if f.endswith(k)), '')
    if '_bench_n' in f:
        desc = "bench.py line under torchrun, N = " + f.split('_bench_n')[1].split('.')[0]
    w(f"* `{f}` — {desc or 'earlier capture of this round, kept for the history of the kernels'}")
print('\n'.join(out))
