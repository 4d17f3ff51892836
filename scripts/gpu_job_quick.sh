#!/bin/bash
# quick check of a change: resident / partition tests, then two short bench lines of the README configuration
T=${1:-r02r}
timeout 400 python -m pytest tests/test_resident.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_tests.txt 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/${T}_tests.txt
for rep in 1 2; do
  timeout 200 python bench.py --workload cfg2 --steps 60 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_ab.json 2>gpurun_out/${T}_ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}_ab.json'))
f=d['roofline']['families_ms']
print('ms/step %.4f e2e %.4f res_pairs %.4f res_combine %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], f['res_pairs'], f['res_combine']))
PY
done
