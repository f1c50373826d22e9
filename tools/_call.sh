mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_step.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
run() { echo "== $1"; env $2 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda 2>>gpurun_out/r2q_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.4f ms  e2e %.4f ms  fused-call %.4f ms  parity %s' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['parity']['ok']))"; }
V=$PWD/svox2_b200/csrc/build/variants
( run nostage "SVOX2_B200_LIB=$V/nostage.so"
  run staged10 ""
  run staged12 "SVOX2_B200_LIB=$V/st12.so"
  run staged8 "SVOX2_B200_LIB=$V/st8.so"
  run staged10_b "" ) 2>&1 | tee gpurun_out/r2q_ab.txt
timeout 120 python tools/timeline.py --out gpurun_out/r2q_timeline.json > gpurun_out/r2q_timeline.txt 2>&1
tail -14 gpurun_out/r2q_timeline.txt
