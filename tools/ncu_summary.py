#!/usr/bin/env python
"""Summarise an `ncu --set full` report of tools/ncu_targets.py into a text table (the metrics B200_PROFILING.md names).
usage: ncu_summary.py report.ncu-rep [label ...] > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max"]


def main():
    rep, labels = sys.argv[1], sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none, one launch per kernel (warm L2: the same launch ran once before)")
    for i, r in enumerate(rows[2:]):
        print(f"\n== launch {i}: {labels[i] if i < len(labels) else ''} :: {r[kn][:70]}")
        for w in WANT:
            if w in h:
                print(f"   {w:72s} {r[h.index(w)]:>18s} {units[h.index(w)]}")


if __name__ == "__main__":
    main()
