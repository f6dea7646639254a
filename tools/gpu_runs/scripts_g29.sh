cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
