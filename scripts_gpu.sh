cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/variant_bench.sh ampa=pnfft_b200/lib/variants/ampa.so
PNFFT_B200_LIB=$PWD/pnfft_b200/lib/variants/ampa.so timeout 600 python -m pytest tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
