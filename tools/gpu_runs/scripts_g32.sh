cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q -k "golden_fixture or family3 or c4_clustered or reference_drivers" > gpurun_out/pytest_f3.log 2>&1; tail -2 gpurun_out/pytest_f3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
