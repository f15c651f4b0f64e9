"""Summarise an `ncu --set full` capture (raw page as CSV) into a markdown table for profiles/.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv > profiles/NAME.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("gpc__cycles_elapsed.avg.per_second", "clock during capture"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active, % of elapsed"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe, % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe active %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 traffic"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld (LDTM) instructions"),
    ("smsp__sass_inst_executed_op_tmem_stt.sum", "tcgen05.st (STTM) instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
names = [r[col["Kernel Name"]].replace("mscs::", "").replace("(int)", "") for r in data]
print("| metric | " + " | ".join(f"`{n.split('(')[0]}`" for n in names) + " |")
print("|---|" + "---|" * len(names))
for key, label in METRICS:
    if key not in col:
        continue
    u = units[col[key]]
    vals = []
    for r in data:
        v = r[col[key]]
        try:
            f = float(v)
            v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.3g}"
        except ValueError:
            pass
        vals.append(v)
    print(f"| {label} ({u}) | " + " | ".join(vals) + " |")
