cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/variant_bench.sh gp2b12=pnfft_b200/lib/variants/gp2b12.so gp2b10=pnfft_b200/lib/variants/gp2b10.so
