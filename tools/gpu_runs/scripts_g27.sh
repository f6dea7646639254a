cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
show() { python - "$1" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms %.2f e2e %.2f val %.3e'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['value']), 'parity', (d.get('parity') or {}).get('parity_rel_l2'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
}
timeout 100 python tools/clustered_bench.py 256 16777216 8 8192 0.05 2>&1 | tail -1
PNFFT_B200_SEG_BALANCE=1 timeout 150 $TR bench.py --gpus $N --config C4 --window gaussian --pre-psi 0 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/b8_c4_bal1.json 2> gpurun_out/b8_c4_bal1.err; show gpurun_out/b8_c4_bal1.json
PNFFT_B200_SEG_BALANCE=0 timeout 150 $TR bench.py --gpus $N --config C4 --window gaussian --pre-psi 0 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/b8_c4_bal0.json 2> gpurun_out/b8_c4_bal0.err; show gpurun_out/b8_c4_bal0.json
