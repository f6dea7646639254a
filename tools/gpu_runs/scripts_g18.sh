cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR bench.py --gpus $N --steps 10 --no-cpu-baseline > gpurun_out/b${N}_final.json 2> gpurun_out/b${N}_final.err
python - $N <<'P'
import json,sys
N=sys.argv[1]
d=json.loads(open("gpurun_out/b%s_final.json"%N).read().strip().splitlines()[-1])
print('N=%s ms %.2f e2e %.2f'%(N,d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'), round(d['roofline']['frac'],3))
for k in ("trafo","adj","trafo_e2e","adj_e2e"): print('   ',k,{a:round(b,2) for a,b in d['stage_ms'][k].items()})
P
