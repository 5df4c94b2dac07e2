"""Compact per-kernel table from `ncu -i report.ncu-rep --page raw --csv` output. Usage: python tools/ncu_table.py raw.csv [every_nth]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("ms", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("blk", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue%", "smsp__issue_active.avg.pct"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
        ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("dramRd", "dram__bytes_read.sum"), ("dramWr", "dram__bytes_write.sum"),
        ("lsb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("ssb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("mathT", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("inst", "smsp__inst_executed.sum")]


def val(r, name):
    if name not in idx:
        return "-"
    v, u = r[idx[name]], units[idx[name]]
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    if name == "gpu__time_duration.sum":
        f *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        return "%.3f" % f
    if "bytes" in name:
        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        return "%.1fMB" % f
    if name == "smsp__inst_executed.sum":
        return "%.2e" % f
    return ("%.1f" % f) if abs(f) < 1e4 else ("%d" % f)


seen = set()
print("| kernel | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|" + "---|" * len(COLS))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
    key = (name, val(r, "launch__grid_size"), val(r, "smsp__inst_executed.sum"))
    if key in seen:
        continue
    seen.add(key)
    print("| " + name + " | " + " | ".join(val(r, m) for _, m in COLS) + " |")
