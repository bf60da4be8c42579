#!/bin/bash
# usage: tools/gpu_variants.sh <size> lib1.so lib2.so ...   (bench value per library variant, strict parity test first)
SZ=$1; shift
for lib in "$@"; do
  echo "== $lib"
  SCG_LIB=$lib timeout 600 python -m pytest tests/test_gpu_md.py -x -q -k "march_matches and mixed3d and morton" 2>&1 | tail -1
  SCG_LIB=$lib timeout 600 python bench.py --workload md3d --size $SZ --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms/substep %.3f' % d['roofline']['avg_launch_ms'])
"
done
