#!/bin/bash
# round-2 GPU session A: environment facts, the GPU test-suite, bench lines for every BASELINE config (CPU arm first so
# that the GPU arm quotes the measured CPU figure), config T last.
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
free -g | head -2; nproc
} > gpurun_out/a_env.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
for c in 2 3 4 5; do
  timeout 600 python bench.py --impl reference --config $c --steps 8 --warmup 3 > gpurun_out/a_ref_cfg$c.json 2> gpurun_out/a_ref_cfg$c.err
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/a_bench_cfg$c.json 2> gpurun_out/a_bench_cfg$c.err
done
( time timeout 900 python bench.py --impl reference --steps 4 --warmup 3 ) > gpurun_out/a_ref_cfgT.json 2> gpurun_out/a_ref_cfgT.err
( time timeout 900 python bench.py --steps 4 --warmup 3 ) > gpurun_out/a_bench_cfgT.json 2> gpurun_out/a_bench_cfgT.err
tail -c 600 gpurun_out/a_pytest.log
for f in gpurun_out/a_bench_cfg*.json gpurun_out/a_ref_cfgT.json; do echo "== $f"; cut -c1-400 $f; done
