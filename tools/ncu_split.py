#!/usr/bin/env python
"""Splits a multi-kernel `ncu --page raw --csv` export into one file per kernel (header + unit row + that kernel's row),
the shape tools/ncu_summary.py and tools/ncu_traffic.py read.  usage: ncu_split.py <raw.csv> <out-prefix>
-> <out-prefix>_<kernel base name>_raw.csv, and a one-line digest per kernel on stdout."""
import csv
import re
import sys


def main():
    raw, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    ki = names.index("Kernel Name")
    seen = set()
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        base = re.sub(r"^void\s+", "", r[ki])
        base = re.split(r"[<(]", base)[0].split("::")[-1].strip()
        if base in seen:
            continue
        seen.add(base)
        with open(f"{prefix}_{base}_raw.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(names); w.writerow(units); w.writerow(r)
        g = lambda k: r[names.index(k)] if k in names else "?"
        print(base, "us", g("gpu__time_duration.sum"), "regs", g("launch__registers_per_thread"), "dram%", g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
              "rd", g("dram__bytes_read.sum"), units[names.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in names else "",
              "wr", g("dram__bytes_write.sum"), "ipc", g("sm__inst_executed.avg.per_cycle_elapsed"), "warps%", g("sm__warps_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
