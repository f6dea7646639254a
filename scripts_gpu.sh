cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -k "c2_full or c3_full or c4_ or c5_" --durations=8 2>&1 | tail -22
