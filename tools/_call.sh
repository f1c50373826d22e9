mkdir -p gpurun_out
run() { echo "== $1"; env $2 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda 2>>gpurun_out/r2t_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%-12s step %.4f ms  e2e %.4f ms  fused-call %.4f ms  parity %s' % ('$1', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['parity']['ok']))"; }
V=$PWD/svox2_b200/csrc/build/variants
( run base ""
  run l2p48 "SVOX2_B200_L2_PERSIST_MB=48"
  run l2p80 "SVOX2_B200_L2_PERSIST_MB=80"
  run l2p32w96 "SVOX2_B200_L2_PERSIST_MB=32 SVOX2_B200_L2_WINDOW_MB=96"
  run tv4 "SVOX2_B200_LIB=$V/tv4.so"
  run tv5 "SVOX2_B200_LIB=$V/tv5.so"
  run base_b "" ) 2>&1 | tee gpurun_out/r2t_ab.txt
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size, 'persist max', getattr(p,'persisting_l2_cache_max_size',None), 'window max', getattr(p,'access_policy_max_window_size',None))
PY
