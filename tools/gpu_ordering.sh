#!/bin/bash
# BASELINE configs[4]: L2 hit rate and DRAM bytes per cell-update of the 3D substep under each cell ordering (ncu, 2e6 cells)
mkdir -p gpurun_out
for o in none morton rcm; do
  timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum \
    --print-units base --clock-control none -k regex:'md_solncfl_staged|md_dsoln_eqo|md_bc_' -s 40 -c 10 --csv --log-file gpurun_out/ordering_$o.csv \
    python bench.py --workload md3d --size 2000000 --order $o --steps 1 --warmup 3 --chunk 1 --no-cpu --no-e2e --no-parity --no-subs > /dev/null 2>&1
done
python - <<'PY'
import csv, collections
print("ordering  kernel                      us      L2hit%  L1hit%  DRAM_MB")
for o in ("none", "morton", "rcm"):
    rows = list(csv.reader(l for l in open(f"gpurun_out/ordering_{o}.csv") if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    idi = hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[idi], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    tot_b, tot_t = 0.0, 0.0
    for (i, k), m in per.items():
        name = k.split("(")[0].replace("void <unnamed>::", "")[:26]
        b = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / 1e6
        print(f"{o:8s}  {name:26s} {m['gpu__time_duration.sum'] / 1e3:8.1f} {m['lts__t_sector_hit_rate.pct']:7.1f} {m['l1tex__t_sector_hit_rate.pct']:7.1f} {b:10.1f}")
        tot_b += b
        tot_t += m['gpu__time_duration.sum'] / 1e3
    print(f"{o:8s}  sum of the captured launches: {tot_t:.0f} us, {tot_b:.0f} MB of DRAM traffic")
PY
