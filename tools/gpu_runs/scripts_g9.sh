cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in gaussian bspline; do for pre in 0 1; do
timeout 300 python bench.py --config C4 --window $w --pre-psi $pre --steps 5 > gpurun_out/bench_c4f_${w}_pre$pre.json 2> gpurun_out/bench_c4f_${w}_pre$pre.err; tail -1 gpurun_out/bench_c4f_${w}_pre$pre.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3), round(d['roofline'].get('gridding_frac'),3))
    except Exception as e: print(f,'ERR',e)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scatter_mma4' -s 3 -c 1 -o gpurun_out/prof_r2_c4s -f python bench.py --config C4 --window gaussian --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_r2_c4s.log 2>&1; tail -2 gpurun_out/ncu_r2_c4s.log | cut -c1-120
