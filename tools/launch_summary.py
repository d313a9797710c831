#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of one
hot-path step (from one vox_insert launch to the next).  Usage: launch_summary.py FILE [step_index]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "vox_insert" in n]
    s, e = starts[step], (starts[step + 1] if step + 1 < len(starts) else len(rows))
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[s:e]:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(n, [0.0, 0])
        a[0] += t
        a[1] += 1
        tot += t
    print(f"# {path}: step {step}, launches {e - s}, total {tot / 1e3:.1f} us ({rows[0]['Metric Unit']} in file)")
    print(f"# {'us':>9} {'share':>6} {'n':>3}  kernel")
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t / 1e3:11.1f} {100 * t / tot:5.1f}% {c:3d}  {n[:110]}")


if __name__ == "__main__":
    main()
