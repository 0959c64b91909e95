#!/bin/bash
# quick GPU visit: parity tests, then per-family step times at 1024^3 for a few env variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
N=${N:-1024}
TAG=new timeout 300 python tools/step_time.py $N 2>&1 | tail -2
for v in "$@"; do
  env $v TAG="$v" timeout 300 python tools/step_time.py $N 2>&1 | tail -2
done
