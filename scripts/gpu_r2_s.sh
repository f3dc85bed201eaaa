#!/bin/bash
# round-2 GPU session S: chain micro-benchmark (final kernels + K-split of the critical update tiles), parity suite,
# cfg2 / cfg3 lines with and without the K-split
mkdir -p gpurun_out
timeout 120 scripts/chain_bench > gpurun_out/s2_chain_bench.log 2>&1; echo "chain_bench rc=$?"
grep -v "marks\|k_tput\|k_lat" gpurun_out/s2_chain_bench.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s2_pytest.log
tail -3 gpurun_out/s2_pytest.log
for ks in 1 4; do
for cfg in 2 3; do
  TLPB200_CRIT_KSPLIT=$ks timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --ipm-device off > gpurun_out/s2_bench_cfg${cfg}_ks$ks.json 2> gpurun_out/s2_bench_cfg${cfg}_ks$ks.err
  echo "cfg$cfg ksplit $ks rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/s2_bench_cfg${cfg}_ks$ks.json').read())
    p=d['phases_one_step']
    print('  value',d['value'],'ms/step',d['ms_per_step'],'update',d['update_ms_host_api'],'diag',p['diag_factor'],'trsm',p['trsm'],'upd',p['update'],'status',d['ipm'].get('status'),d['ipm'].get('iters'))
except Exception as e: print('  parse failed',e)
PY
done
done
