#!/usr/bin/env python
"""Builds profiles/ncu_traffic.json (read by bench.py for roofline.traffic / roofline_fp32) and the short per-kernel
summaries under profiles/ from `ncu --page raw --csv` exports.
usage: ncu_traffic.py <tag> <round-name> stage=raw.csv [stage=raw.csv ...]
  e.g. ncu_traffic.py r02a r02 render_backward=gpurun_out/r02a_render_backward_kernel_raw.csv ..."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(raw):
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, vals = rows[hdr], rows[hdr + 2]
    g = lambda k, d=None: float(vals[names.index(k)].replace(",", "")) if k in names and vals[names.index(k)] not in ("", "n/a") else d
    return g, vals[names.index("Kernel Name")]


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    out = {"workload": "hier3m",
           "how": "ncu --set full --clock-control none --import-source on, one launch per kernel during `bench.py --mode api` (config #3, N=1); "
                  "dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum; issue_slot_util = smsp__inst_executed.sum / "
                  "(SMs x 4 schedulers x sm__cycles_elapsed.max)", "kernels": {}}
    for spec in sys.argv[3:]:
        stage, raw = spec.split("=")
        g, kname = read(raw)
        summary = os.path.join("profiles", f"{rnd}_{stage}_ncu_summary.csv")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw, os.path.join(ROOT, summary), f"{tag}, bench.py --mode api, config #3"],
                       check=True, stdout=subprocess.DEVNULL)
        inst, cyc = g("smsp__inst_executed.sum"), g("sm__cycles_elapsed.max") or g("sm__cycles_elapsed.avg")
        unit_r = 1e6 if True else 1
        dr, dw = g("dram__bytes_read.sum", 0.0), g("dram__bytes_write.sum", 0.0)
        # ncu prints Mbyte / Gbyte / Kbyte depending on size: take the unit row
        rows = list(csv.reader(open(raw)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        units = dict(zip(rows[hdr], rows[hdr + 1]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = dr * scale.get(units.get("dram__bytes_read.sum", "byte"), 1) + dw * scale.get(units.get("dram__bytes_write.sum", "byte"), 1)
        tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        ms = g("gpu__time_duration.sum", 0.0) * tscale.get(units.get("gpu__time_duration.sum", "us"), 1e-3)
        out["kernels"][stage] = {
            "kernel": kname[:80], "dram_bytes": dram, "ms_under_ncu": ms, "warp_inst": inst,
            "ipc": g("sm__inst_executed.avg.per_cycle_elapsed"),
            "issue_slot_util": (inst / (148 * 4 * cyc)) if inst and cyc else None,
            "fma_pipe_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "xu_pipe_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": g("launch__registers_per_thread"), "source": summary}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()
