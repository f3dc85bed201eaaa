#!/bin/bash
# round-2 GPU session G: merged sweeps incl. below items -- GPU test-suite, bench lines, config-5 state-leak diagnostic
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
for c in 3 4 2; do
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_cfg$c.json 2> gpurun_out/g_bench_cfg$c.err
done
timeout 900 python scripts/cfg5_diag.py > gpurun_out/g_cfg5_diag.log 2>&1
tail -c 800 gpurun_out/g_pytest.log
for c in 3 4 2; do echo "== cfg$c"; cut -c1-200 gpurun_out/g_bench_cfg$c.json; done
grep -n "handle\|dc_refine" gpurun_out/g_cfg5_diag.log
