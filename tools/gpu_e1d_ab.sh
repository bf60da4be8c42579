#!/bin/bash
# usage: tools/gpu_e1d_ab.sh lib...  : 1D GPU parity file with the first library, then the 1D Euler bench line per library
first=1
for lib in "$@"; do
  echo "== $lib"
  if [ $first = 1 ]; then SCG_LIB=$lib timeout 600 python -m pytest tests/test_gpu_1d.py tests/test_zz_gpu_golden.py -x -q 2>&1 | tail -2; first=0; fi
  SCG_LIB=$lib timeout 300 python bench.py --workload e1d --size 100000000 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])
"
done
