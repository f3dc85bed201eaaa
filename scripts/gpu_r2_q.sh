#!/bin/bash
# round-2 GPU session R: chain micro-benchmark (phase timeline of k_diag_factor2 / k_trsm2), parity suite, cfg2 / cfg3 lines
mkdir -p gpurun_out
timeout 120 scripts/chain_bench > gpurun_out/r_chain_bench.log 2>&1; echo "chain_bench rc=$?"
grep -v "marks" gpurun_out/r_chain_bench.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r_pytest.log
tail -3 gpurun_out/r_pytest.log
for cfg in 2 3; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --ipm-device off > gpurun_out/r_bench_cfg${cfg}.json 2> gpurun_out/r_bench_cfg${cfg}.err
  echo "cfg$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r_bench_cfg${cfg}.json').read())
    p=d['phases_one_step']
    print('  value',d['value'],'ms/step',d['ms_per_step'],'update',d['update_ms_host_api'],'diag',p['diag_factor'],'trsm',p['trsm'],'upd',p['update'],'status',d['ipm'].get('status'),d['ipm'].get('iters'))
except Exception as e: print('  parse failed',e)
PY
done
