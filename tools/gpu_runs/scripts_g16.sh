cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python bench.py --config C2 --steps 20 > gpurun_out/bench_f_c2.json 2> gpurun_out/bench_f_c2.err; tail -1 gpurun_out/bench_f_c2.err
timeout 500 python bench.py --config C5 --variant c2r --steps 3 > gpurun_out/bench_f_c5_c2r.json 2> gpurun_out/bench_f_c5_c2r.err; tail -1 gpurun_out/bench_f_c5_c2r.err
timeout 500 python bench.py --config C5 --variant float --steps 3 > gpurun_out/bench_f_c5_float.json 2> gpurun_out/bench_f_c5_float.err; tail -1 gpurun_out/bench_f_c5_float.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f val %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu_r2f.log 2>&1; tail -1 gpurun_out/launches_r2f.csv | cut -c1-100
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_(scatter|gather)_mma|k_node_table2' -s 9 -c 3 -o gpurun_out/prof_r2f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_r2f.log 2>&1; tail -1 gpurun_out/ncu_r2f.log | cut -c1-100
