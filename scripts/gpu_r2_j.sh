#!/bin/bash
# round-2 GPU session J (N GPUs): merged-level sweeps inside the sharded phases -- parity (both choreographies), then the bench line
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29581 scripts/dist_check.py mini > gpurun_out/j_dist_mini_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/j_dist_mini_n$N.log
timeout 600 $TR --master-port 29582 scripts/dist_check.py full > gpurun_out/j_dist_full_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/j_dist_full_n$N.log
timeout 1200 $TR --master-port 29583 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err; echo "rc=$?" >> gpurun_out/j_bench_n$N.err
grep -E "DIST-GPU|library" gpurun_out/j_dist_mini_n$N.log gpurun_out/j_dist_full_n$N.log | cut -c1-260
cut -c1-200 gpurun_out/j_bench_n$N.json; tail -2 gpurun_out/j_bench_n$N.err
