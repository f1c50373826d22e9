mkdir -p gpurun_out
timeout 120 python tools/timeline.py --e2e --steps 8 --out gpurun_out/r2s_timeline_e2e.json > gpurun_out/r2s_timeline_e2e.txt 2>&1
tail -22 gpurun_out/r2s_timeline_e2e.txt
