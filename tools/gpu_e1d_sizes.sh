#!/bin/bash
# 1D Euler / 1D scalar rate against problem size: L2-resident sizes take HBM out of the picture
for wl in e1d st1d; do
for sz in 300000 1000000 3000000 10000000 100000000; do
  timeout 300 python bench.py --workload $wl --size $sz --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity --no-subs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', $sz, '%.4g upd/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
done; done
