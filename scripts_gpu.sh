cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/variant_bench.sh kbfast=pnfft_b200/lib/libpnfft_b200.so rpt2=pnfft_b200/lib/variants/rpt2.so
(time timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_gpu.log
