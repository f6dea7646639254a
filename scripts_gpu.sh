cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(scatter|gather)_zm2|k_node_table2' -s 12 -c 4 -f -o gpurun_out/prof_r1b_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
tail -3 gpurun_out/ncu_step.log
