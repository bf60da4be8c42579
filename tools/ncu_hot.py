"""Hottest SASS instructions of a kernel in an ncu source page: share of the stall samples, execution multiplicity, dominant stall.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py <kernel-substring> [top-n] [regions]"""
import csv, sys, collections
want = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 24
nreg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kern, data = None, collections.OrderedDict()
for r in csv.reader(sys.stdin):
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1] if r[1] not in data else None
        continue
    if r and r[0] == "Address":
        if kern is not None:
            data[kern] = (r, [])
        continue
    if kern is not None and kern in data and len(r) > 44 and r[0].startswith("0x"):
        data[kern][1].append(r)
for k, (hdr, rows) in data.items():
    if want not in k:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(x[ix["# Samples"]] or 0) for x in rows)
    base = max(1, int(rows[0][ix["Instructions Executed"]]))
    print("=====", k[:70], "static", len(rows), "samples", tot)
    if nreg:
        n = len(rows)
        tote = sum(int(x[ix["Instructions Executed"]] or 0) for x in rows)
        for b in range(nreg):
            seg = rows[b * n // nreg:(b + 1) * n // nreg]
            s = sum(int(x[ix["# Samples"]] or 0) for x in seg)
            e = sum(int(x[ix["Instructions Executed"]] or 0) for x in seg)
            if not e:
                continue
            st = collections.Counter()
            for x in seg:
                for nm in stalls:
                    st[nm] += int(x[ix[nm]] or 0)
            top = ", ".join(f"{a[6:]} {v / max(s, 1) * 100:.0f}%" for a, v in st.most_common(4))
            print(f"{b * n // nreg:6d}: x{e / len(seg) / base:4.1f} exec {e / tote * 100:5.1f}% samples {s / tot * 100:5.1f}% | {top}")
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]] or 0))[:topn]
    for i in sorted(order):
        x = rows[i]
        s = int(x[ix["# Samples"]] or 0)
        st = {nm[6:]: int(x[ix[nm]] or 0) for nm in stalls}
        print(f"{i:5d} {s / tot * 100:5.2f}% x{int(x[ix['Instructions Executed']]) / base:.1f} {max(st, key=st.get):12s} {x[ix['Source']].strip()[:90]}")
