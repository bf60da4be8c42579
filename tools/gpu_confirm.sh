#!/bin/bash
# usage: tools/gpu_confirm.sh <size3d> <size2d> lib...  : the whole GPU suite with the default library, then per-kernel A/B (3D, 2D)
SZ=$1; SZ2=$2; shift 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
tools/gpu_time.sh $SZ "$@"
echo "==== 2D"
WL=md2d tools/gpu_time.sh $SZ2 "$@"
