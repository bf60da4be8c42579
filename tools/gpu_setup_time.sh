#!/bin/bash
# Setup-time breakdown of the headline case (SCG_HMESH_TIMING=1), no tests.
mkdir -p gpurun_out
SCG_HMESH_TIMING=1 timeout 1200 python bench.py --steps 3 --warmup 3 --no-subs --no-parity --no-cpu --no-e2e "$@" > gpurun_out/bench_setup.json 2> gpurun_out/bench_setup.err
grep "hmesh\]\|dmesh\]" gpurun_out/bench_setup.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_setup.json').read().strip().splitlines()[-1])
print(d['value'], d['config']['mesh']['setup_s'])"
