#!/usr/bin/env python
"""One-line digest of a bench.py JSON line (value, ms/step, parity verdict, per-phase times)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    ex = d.get("exchange") or {}
    print(path, d.get("n_gpus"), d.get("scaling"), round(d.get("value", 0)), d.get("ms_per_step"),
          "e2e", (d.get("e2e") or {}).get("ms_per_step"), "parity", (d.get("parity") or {}).get("ok"),
          ex.get("mode"), ex.get("phase_ms"))
