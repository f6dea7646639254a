cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR tools/mgpu_parity.py > gpurun_out/mgpu$N.log 2>&1; tail -1 gpurun_out/mgpu$N.log | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('mgpu ok',d['ok'],'max', max(max(v['f'],v['grad_f'],v['f_hat']) for v in d['rel_l2'].values()))
except Exception as e: print('mgpu parse ERR', l[-600:])
"
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms %.2f e2e %.2f static %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['e2e_x_static']['ms_per_step']), 'parity', (d.get('parity') or {}).get('parity_rel_l2'), 'launches', d['gpu_launches'], d['library_calls'])
    print('   T',{k:round(v,2) for k,v in d['stage_ms']['trafo'].items()}); print('   A',{k:round(v,2) for k,v in d['stage_ms']['adj'].items()})
except Exception as e: print(sys.argv[1],'ERR',e)
P
}
timeout 400 $TR bench.py --gpus $N --steps 10 > gpurun_out/b${N}_p2p.json 2> gpurun_out/b${N}_p2p.err; show gpurun_out/b${N}_p2p.json; grep -v "OMP_NUM\|^\*\*\*\|^$\|W1017" gpurun_out/b${N}_p2p.err | tail -5
PNFFT_B200_P2P=0 timeout 400 $TR bench.py --gpus $N --steps 10 --no-parity > gpurun_out/b${N}_nccl.json 2> gpurun_out/b${N}_nccl.err; show gpurun_out/b${N}_nccl.json
timeout 400 $TR bench.py --gpus $N --steps 10 --flags 2048 > gpurun_out/b${N}_tr.json 2> gpurun_out/b${N}_tr.err; show gpurun_out/b${N}_tr.json
