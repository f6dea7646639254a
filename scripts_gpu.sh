cd $GRAFT_REPO_ROOT
python tools/clustered_bench.py 256 16777216 6 0 0 | tail -1
python tools/clustered_bench.py 256 16777216 6 0 0.05 | tail -1
python tools/clustered_bench.py 256 16777216 6 0 0.01 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
