cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-parity > gpurun_out/bench_t80.json 2> gpurun_out/bench_t80.err
timeout 200 python bench.py --config C4 --window bspline --steps 3 --no-cpu-baseline --no-parity > gpurun_out/bench_t80_c4b.json 2> gpurun_out/bench_t80_c4b.err
python - <<'P'
import json
for f in ["gpurun_out/bench_t80.json","gpurun_out/bench_t80_c4b.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
P
