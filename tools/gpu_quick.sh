#!/bin/bash
# quick GPU visit: parity tests + bench lines (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
BDR_DEBUG=1 timeout 600 python bench.py --size 512 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench512 rc=$?"
BDR_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "bench1024 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_512.json','gpurun_out/bench_1024.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, 'unreadable', ex); continue
    print(f, 'ms/step %.2f value %.3g e2e %s'%(d['ms_per_step'], d['value'], d['e2e'] and '%.3g'%d['e2e']['value']))
    for k,v in d['kernels'].items(): print('   %-14s %8.3f ms  x%-5.1f %s'%(k, v['ms_per_step'], v['launches_per_step'], ('frac %.3f'%v['frac']) if 'frac' in v else ''))
PY
