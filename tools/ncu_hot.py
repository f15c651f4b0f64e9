"""Top stalled SASS instructions per kernel from an ncu --page source --csv dump.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top] [kernel-substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pick = sys.argv[3] if len(sys.argv) > 3 else ""
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1]
    if pick not in name:
        continue
    hdr = rows[a + 1]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    num = lambda r, k: int(float(r[col[k]] or 0))
    tot = sum(num(r, "# Samples") for r in body)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("==", name, "total samples", tot)
    agg = {s: sum(num(r, s) for r in body) for s in stalls}
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:top]
    for i in sorted(order):
        r = body[i]
        st = {s[6:]: num(r, s) for s in stalls if num(r, s)}
        print(f"{i:5d} {num(r, '# Samples'):7d} {num(r, 'Instructions Executed'):9d}  {r[col['Source']].strip():64s} {st}")
