python scripts/trace_resident.py cfg2 gpurun_out/r02x_unit_times.txt > gpurun_out/r02x_trace_cfg2.txt 2>&1; tail -3 gpurun_out/r02x_trace_cfg2.txt
