set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/r01d_bench_n1.err; tail -c 600 gpurun_out/r01d_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-n200k --no-cpu --no-lbfgs > gpurun_out/r01d_ncu_b2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none -c 120 --csv --log-file gpurun_out/r01d_launches_cfg5_metrics.csv python scripts/prof_run.py cfg5 1 > gpurun_out/r01d_ncu_b5.log 2>&1
bash scripts/ncu_capture.sh cfg5 8 k_potrf_panel r01d_cfg5_panel8
bash scripts/ncu_capture.sh cfg5 0 k_grad r01d_cfg5_grad
ls -la gpurun_out | tail -20
bash scripts/ncu_capture.sh cfg2 0 k_unit_fused r01d_cfg2_fused
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_reference.json 2> gpurun_out/r01d_bench_reference.err
