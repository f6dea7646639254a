cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_tab2.json 2> gpurun_out/bench_tab2.err; tail -2 gpurun_out/bench_tab2.err
for w in gaussian bspline; do
timeout 300 python bench.py --config C4 --window $w --pre-psi 1 --steps 5 --no-cpu-baseline > gpurun_out/bench_c4v4_${w}_pre1.json 2> gpurun_out/bench_c4v4_${w}_pre1.err; tail -2 gpurun_out/bench_c4v4_${w}_pre1.err
done
python - <<'P'
import json
for f in ["gpurun_out/bench_tab2.json","gpurun_out/bench_c4v4_gaussian_pre1.json","gpurun_out/bench_c4v4_bspline_pre1.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, d['roofline']['frac'], d['roofline'].get('gridding_frac'))
    except Exception as e: print(f,'ERR',e)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gather_mma4|k_scatter_zm' -s 6 -c 2 -o gpurun_out/prof_r2_c4 -f python bench.py --config C4 --window gaussian --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_r2_c4.log 2>&1; tail -3 gpurun_out/ncu_r2_c4.log | cut -c1-200
