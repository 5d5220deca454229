#!/usr/bin/env python
"""Aggregates an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`)
by kernel: launches, total time, share. Usage: launch_list_summary.py launches.csv [--last N] [--skip N]
(per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes)."""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)                      # drop the argument list
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("ttb::", "").replace("(anonymous namespace)::", "")
    return name[:90]


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        rows.append((short(r[ki]), v * scale))
    first = int(sys.argv[sys.argv.index("--first") + 1]) if "--first" in sys.argv else 0
    if skip:
        rows = rows[skip:]
    if first:
        rows = rows[:first]
    if last:
        rows = rows[-last:]
    agg = {}
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.1f us total (serialised, cold-cache times under ncu)" % (len(rows), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %6.2f%%  %9.1f us  x%-5d avg %8.2f us  %s" % (100 * a[1] / tot, a[1], a[0], a[1] / a[0], k))


if __name__ == "__main__":
    main()
