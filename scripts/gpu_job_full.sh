#!/bin/bash
# usage: gpu_job_full.sh <tag> : all GPU tests, launch list, short bench
T=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.txt 2>&1; tail -3 gpurun_out/${T}_tests.txt
bash scripts/gpu_job_launches.sh $T | tail -6
python bench.py --steps 20 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'resident', d['roofline']['families_ms']['res_pairs'], 'launches', d['gpu_launches'])
PY
