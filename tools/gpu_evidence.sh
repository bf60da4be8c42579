#!/bin/bash
# ncu evidence of the default 3D kernels for profiles/: launch list of a bench run + one --set full capture (2e6 cells), ordering table
set -u
NSZ=${1:-2000000}
mkdir -p gpurun_out
ARGS="--workload md3d --size $NSZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e --no-parity --no-subs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"md_" -c 400 --csv --log-file gpurun_out/launches_md3d.csv \
   python bench.py $ARGS > gpurun_out/ncu_list_md3d.log 2>&1
python tools/ncu_list.py gpurun_out/launches_md3d.csv | tee gpurun_out/launches_md3d_summary.txt | tail -30
if [ "${LIST_ONLY:-0}" = "1" ]; then exit 0; fi
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"md_solncfl|md_dsoln" -s 12 -c 6 -f -o gpurun_out/prof_md3d \
   python bench.py $ARGS > gpurun_out/ncu_full_md3d.log 2>&1
tail -2 gpurun_out/ncu_full_md3d.log
ncu -i gpurun_out/prof_md3d.ncu-rep --page raw --csv > gpurun_out/prof_md3d_raw.csv
python tools/ncu_sum.py < gpurun_out/prof_md3d_raw.csv > gpurun_out/prof_md3d_summary.txt
grep -A3 "^void" gpurun_out/prof_md3d_summary.txt | grep "void\|time_duration" 
bash tools/gpu_ordering.sh 2>&1 | tee gpurun_out/ordering_ncu.txt | grep "sum of\|ordering"
