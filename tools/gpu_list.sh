#!/bin/bash
# launch list only (cold-cache serialised kernel times) for the md3d bench at a small size
SZ=${1:-2000000}
mkdir -p gpurun_out
ARGS="--workload md3d --size $SZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_md3d.csv python bench.py $ARGS > gpurun_out/ncu_list.log 2>&1
python tools/ncu_list.py gpurun_out/launches_md3d.csv | grep -v "prepare\|fill\|init\|ghost"
