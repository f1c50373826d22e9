#!/usr/bin/env python
"""DRAM traffic of the step's kernels from an `ncu --set full` report -> profiles/r2_traffic.json,
the file bench.py's `roofline.traffic` is read from (never a literal in bench.py).

    tools/ncu_traffic.py gpurun_out/step.ncu-rep c3 [--out profiles/r2_traffic.json]

The report is expected to hold the kernels of (at least) one fused call followed by one whole
train step, as `bench.py --profile --steps 1` brackets them.  For each kernel name the FIRST launch
is used for the fused-call figure; `step_dram_bytes` sums one launch of every kernel of a step
(the fused kernels + TV + optimiser kernels)."""
import argparse
import csv
import json
import os
import re
import subprocess

FUSED = ("trace_kernel", "shade_kernel", "seed_kernel", "scatter_kernel")
STEP_EXTRA = ("tv_", "masked_", "rmsprop", "optim", "compact", "clear_flags", "step_clear", "rank_", "zero_slots")


def short(name):
    m = re.search(r"(\w+)\s*(<[^(]*>)?\s*\(", name)
    return (m.group(1) if m else name)[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("workload")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "profiles", "r2_traffic.json"))
    ap.add_argument("--head", default=None)
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    iname, ird, iwr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    idur = hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}        # -> microseconds
    per = {}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = short(r[iname])
        b = float(r[ird].replace(",", "")) * scale.get(units[ird], 1.0) + \
            float(r[iwr].replace(",", "")) * scale.get(units[iwr], 1.0)
        us = float(r[idur].replace(",", "")) * tscale.get(units[idur], 1.0)
        per.setdefault(k, []).append({"dram_bytes": b, "us": us})
    fused = {k: v[0] for k, v in per.items() if k in FUSED}
    # the train step launches its own variants of the four march kernels (slot-map templates):
    # they are the LAST launch of each name when the capture is bracketed as documented
    step_fused = {k: v[-1] for k, v in per.items() if k in FUSED}
    extra = {k: v for k, v in per.items() if k not in FUSED and any(t in k for t in STEP_EXTRA)}
    head = args.head or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                       text=True).stdout.strip()
    rec = {
        "head": head, "source": os.path.basename(args.report) + ", ncu --set full --clock-control none",
        "fused_call_dram_bytes": sum(v["dram_bytes"] for v in fused.values()),
        "fused_call_kernels": fused,
        # every launch of the non-march kernels in the report belongs to ONE step when the capture
        # was bracketed as documented (TV x2, optimiser x2, ...)
        "step_dram_bytes": sum(v["dram_bytes"] for v in step_fused.values()) +
                           sum(x["dram_bytes"] for v in extra.values() for x in v),
        "step_march_kernels": step_fused,
        "step_other_kernels": {k: v for k, v in extra.items()},
    }
    out = {}
    if os.path.exists(args.out):
        out = json.load(open(args.out))
    out[args.workload] = rec
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
