#!/usr/bin/env python
"""Aggregate the ncu source page (SASS-level stall samples) into windows, to see which part of a
kernel the warps spend their time in.  usage: ncu_regions.py rep.ncu-rep <kernel-regex> [window]"""
import csv
import subprocess
import sys


def toint(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(rep, regex, win=200):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id",
                          f"::regex:{regex}:1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ist = hdr.index("Warp Stall Sampling (All Samples)")
    iex = hdr.index("Instructions Executed")
    isrc = hdr.index("Source")
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    # the page repeats itself once; keep the first copy
    first = data[0][hdr.index("Address")]
    for i in range(1, len(data)):
        if data[i][hdr.index("Address")] == first:
            data = data[:i]
            break
    tot = sum(toint(r[ist]) for r in data)
    print("samples", tot, "sass instructions", len(data), "executed",
          sum(toint(r[iex]) for r in data))
    for k in reasons:
        v = sum(toint(r[hdr.index(k)]) for r in data)
        if v:
            print(f"  {k:28s} {v:8d} {100 * v / tot:5.1f}%")
    for s in range(0, len(data), win):
        seg = data[s:s + win]
        ss = sum(toint(r[ist]) for r in seg)
        ex = sum(toint(r[iex]) for r in seg)
        ops = {}
        for r in seg:
            t = r[isrc].split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
        print(f"[{s:5d}-{s + win:5d}] samples {ss:6d} {100 * ss / max(tot, 1):5.1f}%  exec {ex / 1e6:7.2f}M  {top}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 200)
