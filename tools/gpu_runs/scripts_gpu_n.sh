cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_parity.py > gpurun_out/mgpu$N.log 2>&1; tail -1 gpurun_out/mgpu$N.log | cut -c1-80
