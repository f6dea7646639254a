cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python tools/quick_bench.py 256 16777216 3 6 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -4 gpurun_out/pytest_gpu_h.log
