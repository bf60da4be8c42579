#!/bin/bash
# usage: tools/gpu_prof3.sh <size> <kernel-regex> <skip> <count> <tag> [env...]  -> full capture deep into a normal bench run (chunk 5)
SZ=$1; KRE=$2; SK=$3; CNT=$4; TAG=$5; shift 5
mkdir -p gpurun_out
ARGS="--workload md3d --size $SZ --steps 5 --warmup 3 --chunk 5 --no-cpu --no-e2e"
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SK -c $CNT -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
