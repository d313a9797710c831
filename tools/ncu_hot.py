#!/usr/bin/env python
"""Hottest SASS lines of one kernel from an ncu report's source page.
usage: ncu_hot.py REPORT.ncu-rep LAUNCH_SKIP [TOPN]"""
import csv
import io
import subprocess
import sys


def main():
    rep, skip = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    lines = raw.splitlines()
    print(lines[0][:160])
    ends = [i for i, ln in enumerate(lines) if ln.startswith('"Kernel Name"')]
    end = ends[1] if len(ends) > 1 else len(lines)
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:end]))))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print("total samples", tot, "sass lines", len(rows))
    for i, r in enumerate(rows):
        r["_i"] = i
    top = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:topn]
    for r in sorted(top, key=lambda r: r["_i"]):
        stalls = {k[6:]: int(r[k] or 0) for k in r if k.startswith("stall_") and "Not Issued" not in k and (r[k] or "0") != "0"}
        main_stall = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        print(f"{r['_i']:5d} {100.0 * int(r['# Samples'] or 0) / max(tot, 1):5.1f}% exec={r['Instructions Executed']:>9s} "
              f"shW={r['L1 Wavefronts Shared']:>8s}/{r['L1 Wavefronts Shared Ideal']:>8s} {main_stall}  {r['Source'][:90]}")


if __name__ == "__main__":
    main()
