#!/bin/bash
# usage: tools/gpu_pair.sh <size> <first-pair-lib> lib...  : full md parity file for the first pair library, then tools/gpu_time.sh over all
SZ=$1; FIRST=$2; shift 2
mkdir -p gpurun_out
echo "== full tests/test_gpu_md.py with $FIRST"
SCG_LIB=$FIRST timeout 900 python -m pytest tests/test_gpu_md.py tests/test_gpu_decomp.py -x -q 2>&1 | tail -4
tools/gpu_time.sh $SZ "$@"
