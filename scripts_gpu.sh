cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for d in check_trafo check_adj; do RANK=0 WORLD_SIZE=1 oracle/_ref/drivers/$d -pnfft_np 1 1 1 -pnfft_compute_hessian_f 0 -pnfft_N 16 16 16 2>&1 | tail -12; done
RANK=0 WORLD_SIZE=1 oracle/_ref/drivers/check_vs_pfft -pnfft_np 1 1 1 -pnfft_N 16 16 16 2>&1 | tail -5
RANK=0 WORLD_SIZE=1 oracle/_ref/drivers/pnfft_test 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q -k "drivers or multi" 2>&1 | tail -5
