cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "family3 or c4_clustered" > gpurun_out/pytest_f3.log 2>&1; tail -3 gpurun_out/pytest_f3.log
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0.05 2>&1 | tail -1
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0 2>&1 | tail -1
for w in gaussian bspline; do for pre in 1; do
timeout 300 python bench.py --config C4 --window $w --pre-psi $pre --steps 5 > gpurun_out/bench_c4f_${w}_pre$pre.json 2> gpurun_out/bench_c4f_${w}_pre$pre.err; tail -1 gpurun_out/bench_c4f_${w}_pre$pre.err
done; done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4f_*pre1.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3), round(d['roofline'].get('gridding_frac'),3), d['cpu_baseline']['value'])
    except Exception as e: print(f,'ERR',e)
P
