#!/bin/bash
# 2-GPU check of the no-sync multi-GPU path: ShardedGPRF vs unsharded, a jitter case, then the bench
T=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded.py cfg1 cfg2 cfg5 2>&1 | grep -E "world|Error|error|assert" | head; echo "check rc=${PIPESTATUS[0]}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/check_sharded_jitter.py 2>&1 | grep -E "jitter|Error|error|assert|ok" | head; echo "jitter rc=${PIPESTATUS[0]}"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n2.json'))
print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e'], d['roofline'].get('allreduce_ms'), d['roofline'].get('sync_redos'))
n=d.get('n200k',{}); print('n200k', n.get('value'), n.get('ms_per_step'), n.get('error'))
PY
