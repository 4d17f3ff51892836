#!/bin/bash
T=r02o
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.txt 2>&1; tail -3 gpurun_out/${T}_tests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.err
python scripts/trace_resident.py cfg2 > gpurun_out/${T}_trace_cfg2.txt 2>&1; tail -5 gpurun_out/${T}_trace_cfg2.txt
