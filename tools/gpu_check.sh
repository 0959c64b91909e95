#!/bin/bash
# one GPU-box visit: parity tests, smoke, bench lines, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
BDR_DEBUG=1 timeout 600 python bench.py --size 512 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench512 rc=$?"
BDR_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "bench1024 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_512.csv python bench.py --size 512 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
