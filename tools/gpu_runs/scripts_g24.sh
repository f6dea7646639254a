cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms %.2f e2e %.2f val %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value']), 'parity', (d.get('parity') or {}).get('parity_rel_l2'), round(d['roofline']['frac'],3))
    for k in ("trafo","adj"): print('   ',k,{a:round(b,2) for a,b in d['stage_ms'][k].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
}
timeout 200 $TR bench.py --gpus $N --config C4 --window gaussian --pre-psi 0 --steps 5 --no-cpu-baseline > gpurun_out/b8_c4_pre0.json 2> gpurun_out/b8_c4_pre0.err; show gpurun_out/b8_c4_pre0.json
timeout 200 $TR bench.py --gpus $N --config C4 --window gaussian --pre-psi 1 --steps 5 --no-cpu-baseline > gpurun_out/b8_c4_pre1.json 2> gpurun_out/b8_c4_pre1.err; show gpurun_out/b8_c4_pre1.json
timeout 300 $TR bench.py --gpus $N --config C5 --variant c2r --steps 5 --no-cpu-baseline > gpurun_out/b8_c5_c2r.json 2> gpurun_out/b8_c5_c2r.err; show gpurun_out/b8_c5_c2r.json
