python scripts/golden_diag.py > gpurun_out/r02_golden_runs_device_sampling.txt 2>&1; tail -3 gpurun_out/r02_golden_runs_device_sampling.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02ac_tests.txt 2>&1; tail -4 gpurun_out/r02ac_tests.txt
