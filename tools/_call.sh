mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
run() { echo "== $1"; env $2 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda 2>>gpurun_out/r2p_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.4f ms  e2e %.4f ms  fused-call %.4f ms  parity %s render %.1f' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['parity']['ok'], d['render']['value']))"; }
( run new ""
  run new_b "" ) 2>&1 | tee gpurun_out/r2p_ab.txt
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2p_bench_c2.json 2>gpurun_out/r2p_c2.err; echo c2 rc=$?
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2p_bench_c5.json 2>gpurun_out/r2p_c5.err; echo c5 rc=$?
python tools/show_bench.py gpurun_out/r2p_bench_c2.json gpurun_out/r2p_bench_c5.json
