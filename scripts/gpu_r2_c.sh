#!/bin/bash
# round-2 GPU session C (N GPUs, default 2): sharded solver -- library-issued NCCL inside the CUDA graph vs the phase API,
# parity with the oracle, then the driver's multi-GPU bench line.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c_topo_n$N.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 scripts/dist_check.py mini > gpurun_out/c_dist_mini_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/c_dist_mini_n$N.log
if ! grep -q DIST-GPU-OK gpurun_out/c_dist_mini_n$N.log; then
  TLPB200_DIST_GRAPH=0 timeout 600 $TR --master-port 29552 scripts/dist_check.py mini > gpurun_out/c_dist_mini_nograph_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/c_dist_mini_nograph_n$N.log
fi
timeout 600 $TR --master-port 29553 scripts/dist_check.py full > gpurun_out/c_dist_full_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/c_dist_full_n$N.log
timeout 900 $TR --master-port 29554 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/c_bench_n$N.json 2> gpurun_out/c_bench_n$N.err; echo "rc=$?" >> gpurun_out/c_bench_n$N.err
tail -8 gpurun_out/c_dist_mini_n$N.log; tail -6 gpurun_out/c_dist_full_n$N.log; cut -c1-600 gpurun_out/c_bench_n$N.json; tail -5 gpurun_out/c_bench_n$N.err
