cd $GRAFT_REPO_ROOT
python tools/dbg_zm2.py 2>&1 | grep "^M" 
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/quick_bench.py 256 16777216 1 | tail -3
python tools/quick_bench.py 256 16777216 3 | tail -3
