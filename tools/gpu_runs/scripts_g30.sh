cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 150 $TR bench.py --gpus $N --config C4 --window gaussian --pre-psi 0 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b2_c4.json 2> gpurun_out/b2_c4.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/b2_c4.json").read().strip().splitlines()[-1])
print('N=2 C4 ms %.2f e2e %.2f'%(d['ms_per_step'],d['e2e']['ms_per_step']), d.get('parity'))
P
