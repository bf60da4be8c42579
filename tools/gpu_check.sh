#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list + one full capture of the top kernel.
# usage: tools/gpu_check.sh <workload> <kernel-regex> [extra bench args]
set -u
WL=${1:-e1d}; KRE=${2:-e1d_half_kernel}; shift 2 || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload $WL "$@" > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench rc=$?"
cat gpurun_out/bench_$WL.json; tail -3 gpurun_out/bench_$WL.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$WL.csv \
   python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ncu_list_$WL.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 6 -c 2 -f -o gpurun_out/prof_$WL \
   python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ncu_full_$WL.log 2>&1
ls -la gpurun_out
