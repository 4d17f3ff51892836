#!/bin/bash
# usage: gpu_job_ab.sh <tag> : resident tests + the k_resident time of three bench runs (run-to-run noise)
# every step under its own timeout: a deadlocked kernel must not eat the GPU budget
T=$1
timeout 300 python -m pytest tests -m gpu -x -q -k resident > gpurun_out/${T}_tests.txt 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/${T}_tests.txt
for i in 1 2 3; do
timeout 120 python bench.py --steps 30 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_bench$i.json 2> gpurun_out/${T}_bench.err || { echo "bench rc=$? (timeout = hang)"; break; }
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench$i.json'))
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'resident', d['roofline']['families_ms']['res_pairs'])
PY
done
timeout 120 python scripts/trace_resident.py cfg2 > gpurun_out/${T}_trace_cfg2.txt 2>&1; grep -E "^## |^#   |CTA end" gpurun_out/${T}_trace_cfg2.txt | sed -n '/pair/,$p'
