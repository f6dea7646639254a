cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_parity.py > gpurun_out/mgpu2.log 2>&1; tail -2 gpurun_out/mgpu2.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.json | cut -c1-400
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.4g ms %.2f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']))
"
