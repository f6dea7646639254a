cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
tools/ubench3.bin > gpurun_out/ubench3.log 2>&1
cat gpurun_out/ubench3.log
