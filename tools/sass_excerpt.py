#!/usr/bin/env python
"""profiles/r2_sass_excerpt.txt: per kernel of libsvox2_b200.so, the SASS mnemonics DESIGN.md leans on
(vector reductions, 128-bit loads / stores, atomics, L2 prefetches) counted from `cuobjdump -sass`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "svox2_b200", "csrc", "libsvox2_b200.so")
PAT = re.compile(r"\b(REDG?\.E\.ADD\.F32(?:x4|\.FTZ\.RN)?[.\w]*|RED\.E\.[\w.]+|LDG\.E\.128[.\w]*|STG\.E\.128[.\w]*|"
                 r"ATOMG?\.E\.[\w.]+|CCTL[.\w]*|UTMA\w*|UTCMMA\w*|LDGSTS[.\w]*)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, stats = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
        elif cur:
            for t in PAT.findall(line):
                stats[cur][t] += 1
    out = ["# SASS excerpt of svox2_b200/csrc/libsvox2_b200.so (sm_100a): `cuobjdump -sass`, mnemonics counted per kernel.",
           "# REDG.E.ADD.F32x4 = red.global.add.v4.f32 (vector reduction), LDG.E.128 = 128-bit load, CCTL.E.PF2 = L2 prefetch.",
           "# No UTCMMA / UTMA anywhere: the path has no dense contraction and no tiled bulk copies (DESIGN.md 4).", ""]
    for k, c in stats.items():
        if not c:
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0][:110]
        out.append(name + "\n    " + ", ".join(f"{a} x{b}" for a, b in sorted(c.items())))
    with open(os.path.join(ROOT, "profiles", "r2_sass_excerpt.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(len(stats), "kernels,", sum(1 for c in stats.values() if c), "listed")


if __name__ == "__main__":
    main()
