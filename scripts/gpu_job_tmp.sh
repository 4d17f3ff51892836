python scripts/trace_resident.py cfg2 > gpurun_out/r02aj_trace_cfg2.txt 2>&1; grep -E "units traced|P2|S:" gpurun_out/r02aj_trace_cfg2.txt | tail -16
