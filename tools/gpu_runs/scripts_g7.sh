cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -4 gpurun_out/pytest_gpu_h.log
for w in gaussian; do
timeout 300 python bench.py --config C4 --window $w --steps 5 --no-cpu-baseline > gpurun_out/bench_c4v4_$w.json 2> gpurun_out/bench_c4v4_$w.err; tail -2 gpurun_out/bench_c4v4_$w.err
done
python - <<'P'
import json
for f in ["gpurun_out/bench_c4v4_gaussian.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()})
    except Exception as e: print(f,'ERR',e)
P
