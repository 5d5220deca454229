#!/usr/bin/env python
"""Reads a committed `ncu --set full` capture (.ncu-rep) and records, for the kernel whose name matches --match, the DRAM
traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches) under --key in
profiles/ncu_traffic.json; bench.py copies that number into `roofline.traffic` (it is never typed by hand).

  python tools/ncu_traffic.py --rep profiles/ncu_r02/ar_step.ncu-rep --match ar_step_kernel --key "AR decode step kernel"
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", required=True)
    ap.add_argument("--match", required=True)
    ap.add_argument("--key", required=True)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    need = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"]
    vals = []
    for r in rows[2:]:
        if len(r) != len(hdr) or a.match not in r[col["Kernel Name"]]:
            continue
        vals.append([float(r[col[n]].replace(",", "")) for n in need])
    if not vals:
        raise SystemExit("no launch of a kernel matching %r in %s" % (a.match, a.rep))
    n = len(vals)
    rd, wr, dur = (sum(v[i] for v in vals) / n for i in range(3))
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[a.key] = {"bytes": rd + wr, "read": rd, "write": wr, "duration_ns": dur, "launches": n,
                "source": os.path.relpath(os.path.abspath(a.rep), ROOT)}
    with open(p, "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
    print(a.key, d[a.key])


if __name__ == "__main__":
    main()
