#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, ncu --set full of the resident kernel (cfg2), launch list,
# phase trace, full bench line (N=1), the reference arm, racecheck of the single-CTA bucketing test with the
# resident watchdog relaxed (a launch takes minutes under the tool).
T=${1:-r02}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_gpu_tests.txt 2>&1; tail -3 gpurun_out/${T}_gpu_tests.txt
bash scripts/ncu_capture.sh cfg2 1 k_resident ${T}_ncu_cfg2_resident
cp gprf_b200/csrc/build/gprf_resident_00.o gpurun_out/${T}_res00.o
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_launches.log 2>&1
python scripts/trace_resident.py cfg2 > gpurun_out/${T}_resident_phase_trace_cfg2.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 400 gpurun_out/${T}_bench_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e'], 'launches', d['gpu_launches'])
print(d['roofline'])
print(d.get('lbfgs_full_run'))
for k,v in d.get('configs',{}).items(): print(k, v.get('ms_per_step'), v.get('e2e'), v.get('error'))
print(d['n200k']['ms_per_step'], d['n200k']['roofline']['families_frac_of_peak'])
PY
GPRF_RES_WATCHDOG_S=3000 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_resident.py::test_single_cta_bucketing_equals_radix_sort_path tests/test_resident.py::test_resident_stages -m gpu -x -q > gpurun_out/${T}_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_racecheck.txt | tail -3
