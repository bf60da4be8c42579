#!/bin/bash
# One gpurun call at the end of a round: the GPU suite, then the DEFAULT bench line exactly as the driver runs it
# (--steps 20 --warmup 5: headline + sub_results), then the reference arm.
set -u
mkdir -p gpurun_out
(nproc; free -g | head -2; lscpu | grep -i "model name"; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/host.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
fi
T0=$SECONDS
timeout 2400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? wall $((SECONDS - T0)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  frac %.3f  fp64 %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("roofline_fp64", {}).get("frac")))
print("setup", d["config"]["mesh"]["setup_s"], "parity", d.get("parity"))
for s in d.get("sub_results", []):
    print("  sub", s.get("workload"), "%.4g" % s["value"] if "value" in s else s.get("error"), s.get("roofline", {}).get("frac"), s.get("wall_s"))
PY
tail -3 gpurun_out/bench_default.err
