mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_dist.py -m gpu -x -q -s > gpurun_out/r2r_pytest_4gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2r_pytest_4gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2r_bench_c3_n4_weak.json 2> gpurun_out/r2r_bench_n4.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2r_bench_c3_n4_weak.json
