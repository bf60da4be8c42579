#!/bin/bash
# usage: tools/gpu_prof_lib.sh <lib.so> <size> <kernel-regex> <skip> <count> <tag>  -> one full ncu capture (with source) of a library variant
LIB=$1; SZ=$2; KRE=$3; SK=$4; CNT=$5; TAG=$6
mkdir -p gpurun_out
ARGS="--workload ${WL:-md3d} --size $SZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e --no-parity --no-subs"
SCG_LIB=$LIB timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SK -c $CNT -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_src.csv
python tools/ncu_sum.py < gpurun_out/prof_${TAG}_raw.csv
