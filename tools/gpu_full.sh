#!/bin/bash
# one GPU-box visit: parity tests, smoke, bench line (1024^3), ncu launch list and
# --set full captures of the step's top kernels (512^3, one launch each)
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
BDR_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_1024.json 2> gpurun_out/${TAG}_bench_1024.err; echo "bench1024 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_1024.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
# --set full: first launch of each top kernel in the step at 512^3
for k in k_ongrid_pointers k_resolve k_trace k_edge_bits k_edge_known k_edge_confirm; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^${k}" -c 1 -f -o gpurun_out/${TAG}_${k} python tools/prof_step.py 512 1 > gpurun_out/${TAG}_ncu_${k}.log 2>&1; echo "ncu ${k} rc=$?"
done
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench_1024.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, 'unreadable', ex); continue
    print(f, 'ms/step %.2f value %.3g e2e %s'%(d['ms_per_step'], d['value'], d['e2e'] and '%.3g'%d['e2e']['value']))
    for k,v in d['kernels'].items(): print('   %-14s %8.3f ms  x%-5.1f %s'%(k, v['ms_per_step'], v['launches_per_step'], ('frac %.3f'%v['frac']) if 'frac' in v else ''))
PY
