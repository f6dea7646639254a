cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms %.2f e2e %.2f static %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e_x_static']['ms_per_step']), 'parity', (d.get('parity') or {}).get('parity_rel_l2'), round(d['roofline']['frac'],3))
    for k in ("trafo","adj","trafo_e2e"): print('   ',k,{a:round(b,2) for a,b in d['stage_ms'][k].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
}
timeout 300 $TR tools/mgpu_parity.py > gpurun_out/mgpu$N.log 2>&1; tail -1 gpurun_out/mgpu$N.log | cut -c1-60
timeout 400 $TR bench.py --gpus $N --steps 10 --no-cpu-baseline > gpurun_out/b2_gf.json 2> gpurun_out/b2_gf.err; show gpurun_out/b2_gf.json
