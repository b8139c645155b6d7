#!/bin/bash
# Multi-GPU bench lines (one process per GPU, NCCL), all three workloads.  usage: r2_multigpu.sh N
N=$1
set -x
mkdir -p gpurun_out
for w in c2 c4 c5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $w --no-extra > gpurun_out/mg_${w}_n$N.json 2> gpurun_out/mg_${w}_n$N.err
  tail -c 400 gpurun_out/mg_${w}_n$N.err; head -c 500 gpurun_out/mg_${w}_n$N.json; echo
done
