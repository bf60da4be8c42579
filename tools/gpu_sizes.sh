#!/bin/bash
# per-cell kernel times of the 3D substep against mesh size (is the neighbour state L2-resident at small sizes?)
for sz in 150000 300000 600000 1200000 4000000; do
SCG_TIMING=1 timeout 600 python bench.py --workload md3d --size $sz --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity --no-subs 2>&1 | python -c "
import sys,json,re
t={}
for l in sys.stdin:
    m=re.search(r'timing\] (\S+)\s+\d+ x\s+([\d.]+) us', l)
    if m: t[m.group(1)]=float(m.group(2))
    else:
        try: d=json.loads(l)
        except Exception: continue
n=d['config']['mesh']['ncell']/3.0
print($sz, 'value %.4g' % d['value'], ' '.join('%s %.3f' % (k.replace('solncfl_staged','soln').replace('dsoln_eqo','dsoln'), v*1e3/n) for k,v in sorted(t.items()) if 'solt' not in k), 'ns/cell', d['clocks']['sm_mhz'])"
done
