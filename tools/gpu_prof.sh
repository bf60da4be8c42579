#!/bin/bash
# usage: tools/gpu_prof.sh <size> <kernel-regex> <skip> <count>   -> launch list + full capture
SZ=${1:-2000000}; KRE=${2:-md_}; SK=${3:-0}; CNT=${4:-6}
mkdir -p gpurun_out
ARGS="--workload md3d --size $SZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_md3d.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python tools/ncu_list.py gpurun_out/launches_md3d.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SK -c $CNT -f -o gpurun_out/prof_md3d python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
