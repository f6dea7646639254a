cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
timeout 300 python bench.py --config C4 --window fast_gaussian --steps 5 > gpurun_out/bench_c4f_fast_gaussian_pre0.json 2> gpurun_out/bench_c4f_fast_gaussian_pre0.err; tail -1 gpurun_out/bench_c4f_fast_gaussian_pre0.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4f_fast_gaussian_pre0.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3), round(d['roofline'].get('gridding_frac'),3))
    except Exception as e: print(f,'ERR',e)
P
