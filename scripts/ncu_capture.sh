#!/bin/bash
# usage: ncu_capture.sh <workload> <skip> <kernel-regex> <tag>
# one --set full capture of one launch, exported as raw + source CSV (the .ncu-rep is kept only if small)
wl=$1; skip=$2; rx=$3; tag=$4
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o gpurun_out/$tag \
    python scripts/prof_run.py $wl 2 > gpurun_out/$tag.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --csv > gpurun_out/$tag.source.csv 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page source --print-source cuda --csv > gpurun_out/$tag.cuda.csv 2>/dev/null
ls -la gpurun_out/$tag.ncu-rep
rm -f gpurun_out/$tag.ncu-rep
