cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_final_n1.json").read().strip().splitlines()[-1])
print('ms %.2f e2e %.2f static %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e_x_static']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, d['roofline']['frac'], d['roofline'].get('gridding_frac'), d['cpu_baseline']['value'])
P
