mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -x -q > gpurun_out/r2x_pytest_step.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2x_pytest_step.log
