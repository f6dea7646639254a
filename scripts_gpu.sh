cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# full capture of the two F-only z-march kernels (second iteration = warm)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(scatter|gather)_zm' -s 2 -c 2 -o gpurun_out/prof_zm_f python tools/quick_bench.py 256 16777216 1 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
# launch list of one bench step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_zm.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
