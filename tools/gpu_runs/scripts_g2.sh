cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; tail -3 gpurun_out/pytest_gpu_h.log
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_e2e_new.json 2> gpurun_out/bench_e2e_new.err; tail -2 gpurun_out/bench_e2e_new.err
PNFFT_B200_X_FIRST=0 PNFFT_B200_ADJ_SPECULATE=0 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-parity > gpurun_out/bench_e2e_old.json 2> gpurun_out/bench_e2e_old.err
python - <<'P'
import json
for f in ["gpurun_out/bench_e2e_new.json","gpurun_out/bench_e2e_old.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms %.2f e2e %.2f static %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e_x_static']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'))
        for k in ("trafo_e2e","adj_e2e"): print('   ',k,{a:round(b,2) for a,b in d['stage_ms'][k].items()})
    except Exception as e: print(f,'ERR',e)
P
