mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda > gpurun_out/r2w_bench_c3_n8_weak.json 2> gpurun_out/r2w_bench_n8.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2w_bench_c3_n8_weak.json
