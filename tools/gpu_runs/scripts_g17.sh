cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --config C5 --variant c2r --steps 3 > gpurun_out/bench_f_c5_c2r.json 2> gpurun_out/bench_f_c5_c2r.err; tail -1 gpurun_out/bench_f_c5_c2r.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_f_c5_c2r.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f val %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
P
