#!/bin/bash
# Multi-GPU bench line of config 4 (one process per GPU, NCCL).  usage: r2_multigpu_c4.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --workload c4 --no-extra > gpurun_out/mg9_c4_n$N.json 2> gpurun_out/mg9_c4_n$N.err
tail -c 300 gpurun_out/mg9_c4_n$N.err; head -c 400 gpurun_out/mg9_c4_n$N.json; echo
