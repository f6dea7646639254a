cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -40
python tools/quick_bench.py 256 16777216 1 | tail -3
python tools/quick_bench.py 256 16777216 3 | tail -3
