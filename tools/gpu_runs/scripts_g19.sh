cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "family3 or c4_clustered or golden_fixture" > gpurun_out/pytest_f3.log 2>&1; tail -3 gpurun_out/pytest_f3.log
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0.05 2>&1 | tail -1
timeout 200 python tools/clustered_bench.py 256 16777216 8 0 0 2>&1 | tail -1
timeout 200 python tools/clustered_bench.py 256 16777216 5 0 0 2>&1 | tail -1
PNFFT_B200_GATHER4=1 timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; tail -2 gpurun_out/bench_g4.err
python - <<'P'
import json
for f in ["gpurun_out/bench_g4.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), {k:round(v['ms'],2) for k,v in d['roofline']['kernels'].items()}, d['roofline']['frac'], d['roofline'].get('gridding_frac'))
    except Exception as e: print(f,'ERR',e)
P
