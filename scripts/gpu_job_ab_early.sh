#!/bin/bash
# A/B of the early release of a block's pairs (GPRF_RES_EARLY), then the resident-path tests
T=${1:-r02k}
timeout 300 python -m pytest tests/test_resident.py -m gpu -x -q > gpurun_out/${T}_resident_tests.txt 2>&1; echo "resident tests rc=$?"; tail -2 gpurun_out/${T}_resident_tests.txt
for wl in cfg2 cfg1; do for e in 0 1 0 1; do
  GPRF_RES_EARLY=$e timeout 200 python bench.py --workload $wl --steps 40 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_ab_${wl}_early$e.json 2>gpurun_out/${T}_ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}_ab_${wl}_early$e.json'))
print('$wl early=$e ms/step %.4f e2e %.4f resident %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'].get('resident_path')))
PY
done; done
