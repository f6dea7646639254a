cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/mgpu_parity.py > gpurun_out/mgpu$N.log 2>&1; tail -1 gpurun_out/mgpu$N.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('mgpu ok',d['ok'],'max', max(max(v['f'],v['grad_f'],v['f_hat']) for v in d['rel_l2'].values())); print({k[:40]:max(v['f'],v['grad_f'],v['f_hat']) for k,v in d['rel_l2'].items()})
except Exception as e: print('mgpu parse ERR', l[-800:])
"
timeout 400 $TR bench.py --gpus $N --steps 10 > gpurun_out/b${N}_final.json 2> gpurun_out/b${N}_final.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/b2_final.json").read().strip().splitlines()[-1])
print('N=2 ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), (d.get('parity') or {}).get('parity_rel_l2'))
P
