#!/bin/bash
# round-2 GPU session P: new chain kernels (k_diag_factor2 / k_trsm2): parity suite with the new default, then per-variant
# phase timings on cfg2 / cfg3 (TLPB200_CHAIN_KERNELS = 0 round-1 kernels, 1 no look-ahead, 2 look-ahead)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/p_pytest.log
tail -3 gpurun_out/p_pytest.log
for v in 0 1 2; do
  for cfg in 2 3; do
    TLPB200_CHAIN_KERNELS=$v timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --ipm-device off > gpurun_out/p_bench_cfg${cfg}_v$v.json 2> gpurun_out/p_bench_cfg${cfg}_v$v.err
    echo "cfg$cfg v$v rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/p_bench_cfg${cfg}_v$v.json').read())
    p=d['phases_one_step']
    print('  value',d['value'],'ms/step',d['ms_per_step'],'update',d['update_ms_host_api'],'diag',p['diag_factor'],'trsm',p['trsm'],'upd',p['update'],'status',d['ipm'].get('status'),d['ipm'].get('iters'))
except Exception as e: print('  parse failed',e)
PY
  done
done
