#!/bin/bash
# Round-2 call D (1 GPU): GPU suite, then the headline workload only (no sub_results) for the e2e arm.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 2400 python bench.py --steps 10 --warmup 3 --no-subs ${@} > gpurun_out/bench_nosubs.json 2> gpurun_out/bench_nosubs.err; echo "bench rc=$?"
cut -c1-2500 gpurun_out/bench_nosubs.json; tail -5 gpurun_out/bench_nosubs.err
