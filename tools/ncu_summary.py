#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the columns the roofline discussion uses.

    python tools/ncu_summary.py raw.csv [launches.jsonl] > summary.csv

One row per profiled launch: kernel, grid, block, duration, DRAM bytes read / written, achieved DRAM GB/s and its share of
the measured HBM peak (MEASURED_PEAKS.json), ncu's own dram / sm throughput percentages, issue-slot utilisation,
achieved occupancy, registers, shared-memory bank conflicts, executed warp instructions.  With the JSON lines that
`tools/hbm_kernels.py --once` printed for the same run (one per workload, in launch order of their FIRST kernel) the
algorithmic bytes of the workload are added next to the kernel that dominates it, so traffic / algorithmic is visible.
"""
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read_gb"),
    ("dram__bytes_write.sum", "dram_write_gb"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "ncu_dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "ncu_sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    # the shared-memory port: wavefronts issued by the load/store units (converters, staging, fold) and by the tensor core's
    # operand reads, each as a share of the port's peak over the whole launch
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_lsu_pct"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_tensor_pct"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    while rows and (not rows[0] or rows[0][0] != "ID"):
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    peak = 6543.7
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    have = [(m, n) for m, n in COLS if m in idx]
    w = csv.writer(sys.stdout)
    w.writerow(["id", "kernel", "grid", "block"] + [n for _, n in have] + ["dram_gbs", "dram_frac_of_measured_hbm"])

    def scale(metric, v):
        u = units[idx[metric]].lower()
        v = float(v.replace(",", "")) if v else 0.0
        if metric == "gpu__time_duration.sum":
            return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}.get(u, 1.0)
        if metric.startswith("dram__bytes"):
            return v * {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0, "tbyte": 1e3}.get(u, 1.0)
        return v
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        vals = {n: scale(m, r[idx[m]]) for m, n in have}
        gb = vals.get("dram_read_gb", 0.0) + vals.get("dram_write_gb", 0.0)
        us = vals.get("duration_us", 0.0)
        gbs = gb / (us * 1e-6) if us else 0.0
        name = r[idx["Kernel Name"]]
        name = name.replace("void ", "").replace("<unnamed>::", "")
        name = name.split("(")[0]
        w.writerow([r[idx["ID"]], name, r[idx["Grid Size"]], r[idx["Block Size"]]] +
                   [("%.6g" % vals[n]) for _, n in have] + ["%.1f" % gbs, "%.3f" % (gbs / peak)])


if __name__ == "__main__":
    main()
