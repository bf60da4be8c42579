#!/bin/bash
# Round-2 call A: the whole GPU suite (incl. the reference's own staged tests) without -x, then the default bench
# (stability window: 105 march(5) calls = 525 steps + e2e) and the 2D bench.
set -u
mkdir -p gpurun_out
(nproc; free -g; lscpu | grep -i "model name"; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/host.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -rfE > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_md3d.json 2> gpurun_out/bench_md3d.err; echo "bench rc=$?"
cat gpurun_out/bench_md3d.json; tail -5 gpurun_out/bench_md3d.err
timeout 900 python bench.py --workload md2d --steps 60 --warmup 5 > gpurun_out/bench_md2d.json 2> gpurun_out/bench_md2d.err; echo "bench2d rc=$?"
cat gpurun_out/bench_md2d.json; tail -5 gpurun_out/bench_md2d.err
