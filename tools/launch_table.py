"""Markdown table of the kernels of one step from an ncu launch list (gpu__time_duration.sum CSV).
usage: python tools/launch_table.py gpurun_out/launches.csv > profiles/NAME.md"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
data = [r for r in rows if r and r[0].isdigit()]
names = [r[4] for r in data]
idx = [i for i, n in enumerate(names) if "k_label_hist" in n]
a, b = idx[0], idx[1]
step = data[a:b]
tot = sum(float(r[-1]) for r in step) / 1000
short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("at::vectorized_elementwise_kernel", "ise_kernel")
print("| kernel | grid | block | us | share |\n|---|---|---|---|---|")
for r in step:
    us = float(r[-1]) / 1000
    print(f"| `{short(r[4])}` | {r[8]} | {r[7]} | {us:.1f} | {100*us/tot:.1f}% |")
print(f"| **sum** | | | **{tot:.1f}** | |")
