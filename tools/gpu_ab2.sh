#!/bin/bash
# usage: tools/gpu_ab2.sh <size3d> <size2d> <first-lib> lib...  : full md + decomp parity files for the first library, then per-kernel A/B (3D, then 2D)
SZ=$1; SZ2=$2; FIRST=$3; shift 3
mkdir -p gpurun_out
echo "== full tests/test_gpu_md.py tests/test_gpu_decomp.py with $FIRST"
SCG_LIB=$FIRST timeout 900 python -m pytest tests/test_gpu_md.py tests/test_gpu_decomp.py -x -q 2>&1 | tail -3
tools/gpu_time.sh $SZ "$@"
echo "==== 2D"
WL=md2d tools/gpu_time.sh $SZ2 "$@"
