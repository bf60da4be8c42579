#!/bin/bash
# ncu --set full of the six hot 3D kernels only (2e6 cells), with the source page exported
set -u
NSZ=${1:-2000000}
mkdir -p gpurun_out
ARGS="--workload md3d --size $NSZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e --no-parity --no-subs"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"md_solncfl|md_dsoln" -s 12 -c 6 -f -o gpurun_out/prof_md3d \
   python bench.py $ARGS > gpurun_out/ncu_full_md3d.log 2>&1
tail -2 gpurun_out/ncu_full_md3d.log
ncu -i gpurun_out/prof_md3d.ncu-rep --page raw --csv > gpurun_out/prof_md3d_raw.csv
python tools/ncu_sum.py < gpurun_out/prof_md3d_raw.csv > gpurun_out/prof_md3d_summary.txt
grep -A3 "^void" gpurun_out/prof_md3d_summary.txt | grep "void\|time_duration"
