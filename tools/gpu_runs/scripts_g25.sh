cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for ns in 0 16 32; do
PNFFT_B200_NSEG=$ns timeout 200 python tools/clustered_bench.py 256 16777216 8 8192 0.05 2>&1 | tail -1
done
