python scripts/trace_resident.py cfg2 gpurun_out/r02af_unit_times.txt > gpurun_out/r02af_trace_cfg2.txt 2>&1; tail -2 gpurun_out/r02af_trace_cfg2.txt
