#!/bin/bash
# round-2 GPU session E (N GPUs): the 2-rank pytest, then the driver's multi-GPU bench line (strong config 4 + device-resident loop + weak variant)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/e_pytest_dist_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest_dist_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/e_bench_n$N.json 2> gpurun_out/e_bench_n$N.err; echo "rc=$?" >> gpurun_out/e_bench_n$N.err
tail -3 gpurun_out/e_pytest_dist_n$N.log; cut -c1-300 gpurun_out/e_bench_n$N.json; tail -5 gpurun_out/e_bench_n$N.err
