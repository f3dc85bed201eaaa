#!/bin/bash
# round-2 GPU session W: k_invert_diag2 (block-doubling inverse of the diagonal blocks): GPU test-suite, then cfg4 / cfg3 / cfg2
# lines with the old and the new kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/w_pytest.log
tail -2 gpurun_out/w_pytest.log
for cfg in 4 3 2; do
for v in 0 1; do
  TLPB200_INVERT_KERNEL=$v timeout 600 python bench.py --config $cfg --steps 8 --warmup 3 --no-cpu-baseline --ipm-device off > gpurun_out/w_bench_cfg${cfg}_v$v.json 2> gpurun_out/w_bench_cfg${cfg}_v$v.err
  echo "cfg$cfg invert kernel $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/w_bench_cfg${cfg}_v$v.json').read())
    p=d['phases_one_step']
    print('  value',d['value'],'ms/step',d['ms_per_step'],'update',d['update_ms_host_api'],'solve',d['solve_ms_host_api'],'invert',p['invert_diag'],'status',d['ipm'].get('status'),d['ipm'].get('iters'),d['ipm'].get('pobj'))
except Exception as e: print('  parse failed',e)
PY
done
done
