cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gather_zm2' -s 2 -c 1 -f -o gpurun_out/prof_r1c_gather python tools/quick_bench.py 256 16777216 3 > gpurun_out/ncu_gather.log 2>&1
tail -2 gpurun_out/ncu_gather.log
