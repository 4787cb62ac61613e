#!/bin/bash
# multi-GPU bench lines (torchrun, NCCL): weak-scaling C2 and strong-scaling C4 at N = $1
N=$1
mkdir -p gpurun_out
for cfg in C2 C4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --config $cfg > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err
  echo "== $cfg N=$N rc=$?"; cut -c1-700 gpurun_out/bench_${cfg}_n$N.json; tail -n 3 gpurun_out/bench_${cfg}_n$N.err
done
