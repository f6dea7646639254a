cd $GRAFT_REPO_ROOT
for cf in 1 3; do timeout 300 python tools/quick_bench.py 256 16777216 $cf 6 0 0 2>&1 | tail -3; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
