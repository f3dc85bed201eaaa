#!/bin/bash
# round-2 GPU session O: last check of the final code -- smoke(), the GPU test-suite, a short config-T bench line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/o_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
( time timeout 900 python bench.py --steps 4 --warmup 3 --ipm-limit 0 ) > gpurun_out/o_bench_cfgT.json 2> gpurun_out/o_bench_cfgT.err
tail -2 gpurun_out/o_smoke.log; tail -c 300 gpurun_out/o_pytest.log; cut -c1-250 gpurun_out/o_bench_cfgT.json; tail -4 gpurun_out/o_bench_cfgT.err
