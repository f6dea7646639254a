cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "family3" > gpurun_out/pytest_f3.log 2>&1; tail -12 gpurun_out/pytest_f3.log
timeout 200 python tools/clustered_bench.py 128 2000000 8 0 0.05 2>&1 | tail -2
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0.05 2>&1 | tail -2
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -5 gpurun_out/pytest_gpu_h.log
