#!/bin/bash
# usage: gpu_job_quick.sh <tag> [pytest -k expression]  - resident tests, phase trace, short bench (cfg2 only)
T=$1; K=${2:-resident}
python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${T}_tests.txt 2>&1; tail -3 gpurun_out/${T}_tests.txt
python scripts/trace_resident.py cfg2 > gpurun_out/${T}_trace_cfg2.txt 2>&1; grep -E "^## |^#   P|finalize|CTA end" gpurun_out/${T}_trace_cfg2.txt
python bench.py --steps 20 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['families_ms'])
PY
