#!/bin/bash
# Round-2 call C (N GPUs): the default bench under torchrun: NCCL parity leg, weak-scaling headline, strong-scaling sub-result.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
(nproc; free -g) > gpurun_out/host_$N.txt 2>&1
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
cut -c1-4000 gpurun_out/bench_n$N.json; tail -30 gpurun_out/bench_n$N.err | head -20
