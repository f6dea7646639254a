cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
timeout 200 python tools/clustered_bench.py 256 16777216 8 8192 0.05 2>&1 | tail -1
timeout 200 python tools/clustered_bench.py 256 16777216 6 0 0.05 2>&1 | tail -1
