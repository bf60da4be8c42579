#!/bin/bash
# usage: tools/gpu_var4.sh <size> <kernel-grep> lib...  : md GPU tests (full file for the first lib, subset for the rest), bench, per-kernel launch times
SZ=$1; KG=$2; shift 2
mkdir -p gpurun_out
first=1
for lib in "$@"; do
  echo "== $lib"
  if [ $first = 1 ]; then SCG_LIB=$lib timeout 1200 python -m pytest tests/test_gpu_md.py -x -q 2>&1 | tail -2; first=0
  else SCG_LIB=$lib timeout 900 python -m pytest tests/test_gpu_md.py -x -q -k "march_matches and (mixed3d or mixed2d) and morton and 3-" 2>&1 | tail -1; fi
  SCG_LIB=$lib timeout 600 python bench.py --workload md3d --size $SZ --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('value %.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms/substep %.3f' % d['roofline']['avg_launch_ms'])
"
  SCG_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/l_$lib.csv python bench.py --workload md3d --size 2000000 --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e > /dev/null 2>&1
  python tools/ncu_list.py gpurun_out/l_$lib.csv | grep "$KG"
done
