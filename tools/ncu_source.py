#!/usr/bin/env python
"""Per-source-line digest of one kernel of an .ncu-rep captured with --import-source on: warp
instructions executed, stall samples and divergence per CUDA source line (aggregated over the SASS
lines ncu attributes to it).  usage: tools/ncu_source.py rep kernel_regex [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(rep, kern, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern,
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    agg, src = defaultdict(lambda: [0, 0, 0, 0]), {}
    seen, fname = set(), ""
    for r in rows:
        if r and r[0] == "File Path":       # one section per source file; a repeated file = the next launch
            if r[1] in seen:
                break
            seen.add(r[1])
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if r and r[0] == "Function Name":
            if len(seen) == 1:
                print("==", r[1][:120])
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            i_thr, i_div = hdr.index("Thread Instructions Executed"), hdr.index("Divergent Branches")
            continue
        if not seen or len(r) < 10 or not r[0].isdigit():
            continue
        line = (fname, int(r[0]))
        src.setdefault(line, r[1])
        a = agg[line]
        for k, i in enumerate((i_inst, i_samp, i_thr, i_div)):
            try:
                a[k] += int(float(r[i] or 0))
            except ValueError:
                pass
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%22s %7s %7s %6s %5s  %s" % ("line", "inst%", "samp%", "thr/in", "div", "source"))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%22s %6.1f%% %6.1f%% %6.1f %5d  %s" % ("%s:%d" % line, 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s,
                                                    a[2] / max(a[0], 1), a[3], src[line].strip()[:100]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
