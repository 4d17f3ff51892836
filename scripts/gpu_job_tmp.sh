python -m pytest tests -m gpu -x -q -k "partition or reblock or bucketing or resident or cfg1 or seismic or sharded" > gpurun_out/r02ap_tests.txt 2>&1; tail -2 gpurun_out/r02ap_tests.txt
bash scripts/gpu_job_launches.sh r02ap | tail -5
python bench.py --steps 30 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/r02ap_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02ap_bench.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['families_ms']['res_pairs'])"
