#!/bin/bash
# bench.py under torchrun on N GPUs of one box (gpurun --gpus N -- bash tools/gpu_scale.sh TAG N)
TAG=$1; N=$2
mkdir -p gpurun_out
BDR_DEBUG=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', d['config']['workload'][:24], 'ms/step %.2f value %.3g e2e %.3g'%(d['ms_per_step'], d['value'], d['e2e']['value']))
PY
