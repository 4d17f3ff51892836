set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err; tail -c 300 gpurun_out/r01f_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01f_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-n200k --no-cpu --no-lbfgs > gpurun_out/r01f_ncu_b2.log 2>&1
bash scripts/ncu_capture.sh cfg2 2 k_potrf_panel r01f_cfg2_panel2
