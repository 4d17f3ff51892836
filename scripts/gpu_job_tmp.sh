timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02au_tests.txt 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02au_tests.txt
for i in 1 2; do timeout 120 python bench.py --steps 30 --warmup 5 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/r02au_bench.json 2>/dev/null || echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02au_bench.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['families_ms']['res_pairs'])"; done
