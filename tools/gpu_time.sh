#!/bin/bash
# usage: tools/gpu_time.sh <size> lib...  : quick parity subset + bench with in-run per-kernel event timing (SCG_TIMING=1)
SZ=$1; shift
for lib in "$@"; do
  echo "== $lib"
  SCG_LIB=$lib timeout 900 python -m pytest tests/test_gpu_md.py -x -q -k "march_matches and (mixed3d or mixed2d or step2d) and morton and 3-" 2>&1 | tail -1
  SCG_LIB=$lib SCG_TIMING=1 timeout 600 python bench.py --workload ${WL:-md3d} --size $SZ --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception:
        if 'timing' in l and ('solncfl' in l or 'dsoln' in l or 'solt' in l): print(l.rstrip())
        continue
    print('value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms/substep %.3f' % d['roofline']['avg_launch_ms'])
"
done
