cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_parity.py > gpurun_out/mgpu2.log 2>&1; grep -v "^\s*$" gpurun_out/mgpu2.log | grep -i "nccl\|pnfft\|{" | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1200 gpurun_out/bench_n2.json; grep -i "nccl\|pnfft" gpurun_out/bench_n2.err | head
