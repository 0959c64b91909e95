#!/bin/bash
# ncu --set full capture of the first launch of each named kernel in one step at N^3 (default 512)
# usage: tools/gpu_ncu.sh TAG kernel [kernel...]
TAG=$1; shift
N=${N:-512}
mkdir -p gpurun_out
for k in "$@"; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^${k}" -c ${C:-1} -f -o gpurun_out/${TAG}_${k} python tools/prof_step.py $N 1 > gpurun_out/${TAG}_ncu_${k}.log 2>&1; echo "ncu ${k} rc=$?"
done
