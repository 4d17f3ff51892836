python -m pytest tests -m gpu -x -q -k "resident or parity_small or cfg1" > gpurun_out/r02ah_tests.txt 2>&1; tail -2 gpurun_out/r02ah_tests.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_resident -s 1 -c 2 python scripts/prof_run.py cfg2 3 2>&1 | grep -E "dram__|gpu__time|k_resident" | head -12
python bench.py --steps 30 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/r02ah_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02ah_bench.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['families_ms']['res_pairs'])"
