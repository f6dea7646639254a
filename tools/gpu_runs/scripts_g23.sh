cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "column_batches or family3" > gpurun_out/pytest_f3.log 2>&1; tail -5 gpurun_out/pytest_f3.log
