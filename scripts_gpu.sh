cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_zm2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(scatter|gather)_zm2|k_node_table2' -s 12 -c 4 -o gpurun_out/prof_zm2_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
