#!/bin/bash
# round-2 GPU session H (N GPUs): the driver's multi-GPU bench line after the launch-skipping / merged-sweep changes
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/h_bench_n$N.json 2> gpurun_out/h_bench_n$N.err; echo "rc=$?" >> gpurun_out/h_bench_n$N.err
cut -c1-300 gpurun_out/h_bench_n$N.json; tail -3 gpurun_out/h_bench_n$N.err
