#!/bin/bash
# quick loop: md GPU tests + resident-only bench lines for a few sizes/modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_md.py -x -q 2>&1 | tail -4
for mode in 2 1; do
SCG_MD_FUSED=$mode timeout 600 python bench.py --workload md3d --size ${1:-4000000} --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('mode $mode', 'value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms/substep %.3f' % d['roofline']['avg_launch_ms'], d['config']['mesh']['setup_s'])
"
done
