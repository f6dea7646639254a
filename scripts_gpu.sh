cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_zm2' -s 1 -c 1 -o gpurun_out/prof_zm2d_g python tools/quick_bench.py 256 16777216 3 > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/ncu_f2.log
