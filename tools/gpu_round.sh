#!/bin/bash
# One gpurun call: GPU parity tests, the 50M-cell bench line, launch list + full ncu capture of the two hot kernels.
# usage: tools/gpu_round.sh <bench-size> <ncu-size> [extra bench args]
set -u
SZ=${1:-50000000}; NSZ=${2:-2000000}; shift 2 || true
mkdir -p gpurun_out
(nproc; free -g; lscpu | grep -i "model name"; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --workload md3d --size $SZ "$@" > gpurun_out/bench_md3d.json 2> gpurun_out/bench_md3d.err; echo "bench rc=$?"
cat gpurun_out/bench_md3d.json; tail -5 gpurun_out/bench_md3d.err
ARGS="--workload md3d --size $NSZ --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_md3d.csv \
   python bench.py $ARGS "$@" > gpurun_out/ncu_list_md3d.log 2>&1
python tools/ncu_list.py gpurun_out/launches_md3d.csv | grep -v "prepare\|fill\|init\|ghost"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"md_solncfl|md_dsoln" -s 12 -c 6 -f -o gpurun_out/prof_md3d \
   python bench.py $ARGS "$@" > gpurun_out/ncu_full_md3d.log 2>&1
tail -3 gpurun_out/ncu_full_md3d.log
cat gpurun_out/host.txt
