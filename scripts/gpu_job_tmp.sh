python scripts/trace_resident.py cfg2 > gpurun_out/r02q_trace_cfg2.txt 2>&1; grep -E "^## |^#   |CTA end" gpurun_out/r02q_trace_cfg2.txt
