"""Warp-stall shares per kernel (and the most-sampled SASS instructions) from an `ncu --page source --csv` dump.
usage: python tools/stall_summary.py gpurun_out/prof_final_src.csv [kernel-substring ...] > profiles/NAME.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
picks = sys.argv[2:] or ["k_sim_fwd", "k_sim_bwd"]
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1]
    if not any(p in name for p in picks) or name in seen:
        continue
    seen.add(name)
    hdr = rows[a + 1]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    num = lambda r, k: int(float(r[col[k]] or 0))
    tot = sum(num(r, "# Samples") for r in body)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(num(r, s) for r in body), s[6:]) for s in stalls), reverse=True)
    print(f"## `{name}` -- {tot} samples, {len(body)} SASS instructions\n")
    print("| stall reason | share |\n|---|---|")
    for v, s in agg:
        if v * 100 >= tot:
            print(f"| {s} | {100.0 * v / tot:.1f}% |")
    print("\n| # | samples | executed | instruction | main reasons |\n|---|---|---|---|---|")
    for i in sorted(sorted(range(len(body)), key=lambda i: -num(body[i], "# Samples"))[:12]):
        r = body[i]
        st = sorted(((num(r, s), s[6:]) for s in stalls if num(r, s)), reverse=True)[:2]
        print(f"| {i} | {num(r, '# Samples')} | {num(r, 'Instructions Executed')} | `{r[col['Source']].strip()}` | "
              + ", ".join(f"{s} {v}" for v, s in st) + " |")
    print()
