cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_parity.py > gpurun_out/mgpu$N.log 2>&1; tail -1 gpurun_out/mgpu$N.log | cut -c1-60
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value %.4g ms %.2f e2e %.4g'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms']['trafo'].items()}, {k:round(v,2) for k,v in d['stage_ms']['adj'].items()})
"
