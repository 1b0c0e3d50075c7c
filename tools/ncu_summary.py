#!/usr/bin/env python
"""Compact per-kernel summary of `ncu -i X.ncu-rep --page raw --csv` output (one markdown table row per kernel)."""
import csv, sys
KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("fovgs::", "")
        out = [name]
        for k, lab in KEYS:
            if k in h:
                i = h.index(k)
                out.append(f"{lab}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
        st = [(float(r[i]), h[i]) for i in range(len(h)) if "issue_stalled" in h[i] and h[i].endswith("_per_issue_active.ratio") and "not_issued" not in h[i] and r[i]]
        top = ", ".join("%s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, n in sorted(st, reverse=True)[:4])
        print("| " + " | ".join(out) + " | stalls/issue: " + top + " |")
