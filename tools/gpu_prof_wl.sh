#!/bin/bash
# usage: tools/gpu_prof_wl.sh <workload> <size> <kernel-regex> <skip> <count> <tag>  -> launch list + one full capture for any bench workload
WL=$1; SZ=$2; KRE=$3; SK=$4; CNT=$5; TAG=$6
mkdir -p gpurun_out
ARGS="--workload $WL --size $SZ --steps 1 --warmup 3 --chunk 2 --no-cpu --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > /dev/null 2>&1
python tools/ncu_list.py gpurun_out/launches_$TAG.csv | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SK -c $CNT -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/ncu_full_$TAG.log
