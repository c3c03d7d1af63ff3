#!/usr/bin/env python
"""ncu `--page raw --csv` export -> the short `metric,value,unit` table kept under profiles/.
usage: ncu_summary.py <raw.csv> <out.csv> ["note for the header line"]"""
import csv
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    kname = vals[names.index("Kernel Name")]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, {kname[:90]}, {note}\n")
        for m in KEEP:
            if m in names:
                i = names.index(m)
                v = vals[i].replace(",", "")
                try:
                    v = f"{float(v):.6f}"
                except ValueError:
                    pass
                f.write(f"{m},{v},{units[i]}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
