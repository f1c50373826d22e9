#!/usr/bin/env bash
# Run bench.py (kernel-only numbers) for the default library and every build variant.
cd "$(dirname "$0")/.."
run() { SVOX2_B200_LIB="$2" python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-ref-cuda ${WL:+--workload $WL} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.3f ms  e2e %.3f ms  fused-call %.3f ms  frac %.3f  render %.1f MPix/s' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['render']['value']))"; }
run default ""
shopt -s nullglob
for so in svox2_b200/csrc/build/variants/*.so; do run "$(basename $so .so)" "$PWD/$so"; done
