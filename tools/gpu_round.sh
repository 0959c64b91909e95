#!/bin/bash
# evidence run for profiles/: tests, smoke, bench lines (ours + reference arm), launch list,
# DRAM traffic of every launch of one step, --set full captures of the top kernels
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -rs > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
BDR_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_1024.json 2> gpurun_out/${TAG}_bench_1024.err; echo "bench1024 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "benchref rc=$?"
for w in c3 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w rc=$?"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches_1024.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'^k_(?!synth)' -c 400 --csv --log-file gpurun_out/${TAG}_traffic_1024.csv python tools/prof_step.py 1024 1 > gpurun_out/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
N=1024 bash tools/gpu_ncu.sh ${TAG} k_trace k_seed_pointers k_resolve_tiles k_label_eq_bits k_edge_known
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
