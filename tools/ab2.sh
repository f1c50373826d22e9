cd /root/repo
run() { SVOX2_B200_LIB="$2" python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-ref-cuda --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.4f ms  e2e %.4f ms' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step']))"; }
for i in 1 2 3; do
run default ""
run zopt "$PWD/svox2_b200/csrc/build/variants/zopt.so"
run b16 "$PWD/svox2_b200/csrc/build/variants/b16.so"
done
