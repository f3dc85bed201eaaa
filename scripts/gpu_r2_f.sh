#!/bin/bash
# round-2 GPU session F: merged-level sweeps -- GPU test-suite, config 3 / 4 / 2 bench lines, config-5 host-vs-device IPM diagnostic
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
if ! grep -q "pytest rc=0" gpurun_out/f_pytest.log; then
  TLPB200_MERGE_LEVELS=0 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest_nomerge.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_nomerge.log
fi
for c in 3 4 2; do
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_cfg$c.json 2> gpurun_out/f_bench_cfg$c.err
done
timeout 600 python scripts/cfg5_diag.py > gpurun_out/f_cfg5_diag.log 2>&1
TLPB200_DC_REFINE=4 timeout 600 python scripts/cfg5_diag.py > gpurun_out/f_cfg5_diag_refine4.log 2>&1
tail -c 1500 gpurun_out/f_pytest.log
for c in 3 4 2; do echo "== cfg$c"; cut -c1-200 gpurun_out/f_bench_cfg$c.json; done
head -3 gpurun_out/f_cfg5_diag.log; head -3 gpurun_out/f_cfg5_diag_refine4.log
