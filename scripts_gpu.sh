cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "window_tensor and kaiser" 2>&1 | tail -1
