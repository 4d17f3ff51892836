#!/bin/bash
# A/B of the resident path's scheduling switches on the README configuration (device step and e2e), interleaved
T=${1:-r02n}
timeout 300 python -m pytest tests/test_resident.py -m gpu -x -q > gpurun_out/${T}_resident_tests.txt 2>&1; echo "resident tests rc=$?"; tail -1 gpurun_out/${T}_resident_tests.txt
for rep in 1 2 3; do for cfg in "0 0 0" "1 0 0" "1 1 0" "1 0 1" "1 1 1"; do
  set -- $cfg
  GPRF_RES_EARLY=$1 GPRF_RES_DEFER=$2 GPRF_RES_SORTBLK=$3 timeout 200 python bench.py --workload cfg2 --steps 60 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_ab.json 2>gpurun_out/${T}_ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}_ab.json'))
print('early=$1 defer=$2 sortblk=$3  ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done; done
