#!/bin/bash
# usage: gpu_job_n.sh <tag> <N> : bench at N GPUs (torchrun as the driver launches it)
T=$1; N=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
tail -c 300 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n$N.json'))
print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline'].get('allreduce_ms'))
n=d.get('n200k',{})
print('n200k', n.get('value'), n.get('ms_per_step'), n.get('error'))
PY
