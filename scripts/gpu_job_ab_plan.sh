#!/bin/bash
# A/B of the split bucketing launch on the README configuration (device step and e2e), interleaved
T=${1:-r02q}
timeout 400 python -m pytest tests/test_resident.py tests/test_gpu_parity.py -m gpu -x -q -k "resident or partition or bucket or tree or reblock or update_X" > gpurun_out/${T}_tests.txt 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/${T}_tests.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_resident.py::test_single_cta_bucketing_equals_radix_sort_path" -m gpu -x -q > gpurun_out/${T}_memcheck.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_memcheck.txt | tail -2
for rep in 1 2 3; do for sp in 0 4 2 8; do
  GPRF_BUCKET_SPLIT=$sp timeout 200 python bench.py --workload cfg2 --steps 60 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_ab.json 2>gpurun_out/${T}_ab.err
  python - <<PY
import json
d=json.load(open('gpurun_out/${T}_ab.json'))
print('split=$sp  ms/step %.4f e2e %.4f' % (d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done; done
