mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2u_bench_c3.json 2> gpurun_out/r2u_bench_c3.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2u_bench_c3.json
