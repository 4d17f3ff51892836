#!/bin/bash
# usage: gpu_job_launches.sh <tag> : ncu launch list (durations) of a short cfg2 bench run
T=$1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-n200k --no-cpu --no-extra --no-lbfgs > gpurun_out/${T}_launches.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${T}_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:60].ljust(60), r[-1], r[-2])
PY
