cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-220
