#!/bin/bash
# usage: gpu_job_ncu.sh <tag> <workload> <kernel regex> <skip>
T=$1; WL=${2:-cfg2}; RX=${3:-k_resident}; SKIP=${4:-2}
bash scripts/ncu_capture.sh $WL $SKIP "$RX" ${T}_ncu_${WL}
cp gprf_b200/csrc/build/gprf_resident_00.o gpurun_out/${T}_res00.o
