#!/bin/bash
# usage: gpu_job_tests.sh <tag> <pytest args...>
T=$1; shift
python -m pytest "$@" > gpurun_out/${T}_tests.txt 2>&1; tail -40 gpurun_out/${T}_tests.txt
