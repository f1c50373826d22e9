mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
run() { echo "== $1"; env $2 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda 2>>gpurun_out/r2m_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.4f ms  e2e %.4f ms  fused-call %.4f ms  parity %s' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['parity']['ok']))"; }
V=$PWD/svox2_b200/csrc/build/variants
( run head "SVOX2_B200_LIB=$V/head.so"
  run new ""
  run ring2 "SVOX2_B200_LIB=$V/ring2.so"
  run new_b "" ) 2>&1 | tee gpurun_out/r2m_ab.txt
timeout 120 python tools/timeline.py --out gpurun_out/r2m_timeline.json > gpurun_out/r2m_timeline.txt 2>&1
tail -14 gpurun_out/r2m_timeline.txt
