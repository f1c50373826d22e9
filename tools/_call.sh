mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"masked|tv_|scatter|shade|seed|trace|rmsprop|rank_|step_clear|clear_flags|lattice" -c 22 -o gpurun_out/r2_final_step python bench.py --steps 1 --warmup 3 --profile --profile-steps 1 --no-cpu-baseline --no-ref-cuda --no-parity > gpurun_out/r2_final_ncu.log 2>&1; echo ncu rc=$?
python tools/ncu_traffic.py gpurun_out/r2_final_step.ncu-rep c3 --head 063e755 > gpurun_out/r2_final_traffic.log 2>&1; cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --profile --no-cpu-baseline --no-ref-cuda --no-parity > gpurun_out/r2_final_launch.log 2>&1; echo launches rc=$?
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest_1gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_1gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final_smoke.log
timeout 400 python bench.py > gpurun_out/r2_final_bench_c3.json 2> gpurun_out/r2_final_bench_c3.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2_final_bench_c3.json
