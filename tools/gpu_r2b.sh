#!/bin/bash
# Round-2 call B (1 GPU): full GPU suite with -x like the driver, then the default bench with sub_results.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 2400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/bench_default.json; tail -25 gpurun_out/bench_default.err | head -12
