"""Summarise an ncu source page (csv on stdin): per code region, executed share, sample share and the stall mix.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_src.py [kernel-substring] [nbuckets]"""
import csv, sys, collections
want = sys.argv[1] if len(sys.argv) > 1 else ""
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rd = csv.reader(sys.stdin)
kern, hdr, rows, take = None, None, [], False
for r in rd:
    if len(r) >= 2 and r[0] == "Kernel Name":
        if take and rows:
            break
        kern = r[1]
        take = want in kern
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if take and len(r) > 44 and r[0].startswith("0x"):
        rows.append(r)
names = [h for h in hdr[27:44]]
n = len(rows)
tot = sum(int(x[2] or 0) for x in rows)
tote = sum(int(x[5] or 0) for x in rows)
print(kern, "static", n, "exec", tote, "samples", tot)
for b in range(B):
    seg = rows[b * n // B:(b + 1) * n // B]
    s = sum(int(x[2] or 0) for x in seg)
    e = sum(int(x[5] or 0) for x in seg)
    st = collections.Counter()
    for x in seg:
        for i, nm in enumerate(names):
            st[nm] += int(x[27 + i] or 0)
    top = ", ".join(f"{k[6:]} {v / max(s, 1) * 100:.0f}%" for k, v in st.most_common(5))
    print(f"{b * n // B:6d}: exec {e / tote * 100:5.1f}% samples {s / tot * 100:5.1f}% | {top}")
