#!/bin/bash
# usage: tools/gpu_staged.sh <size> <modes> lib1.so ...  : parity (march_matches, mixed3d+hex3d+mixed2d, all orders for given modes) then bench per lib x mode
SZ=$1; MODES=$2; shift 2
mkdir -p gpurun_out
for lib in "$@"; do
  echo "== $lib"
  SCG_LIB=$lib timeout 900 python -m pytest tests/test_gpu_md.py -x -q -k "march_matches and (mixed3d or prism3d or mixed2d) and morton" 2>&1 | tail -2
  for mode in $MODES; do
  SCG_LIB=$lib SCG_MD_FUSED=$mode timeout 600 python bench.py --workload md3d --size $SZ --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('mode $mode value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms/substep %.3f' % d['roofline']['avg_launch_ms'])
"
  done
done
