#!/bin/bash
# round-2 GPU session M: re-validation after the wait re-ordering / per-phase refactor
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
for c in 3 4 2; do
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_cfg$c.json 2> gpurun_out/m_bench_cfg$c.err
done
tail -c 300 gpurun_out/m_pytest.log
for c in 3 4 2; do python - <<PY
import json
d=json.load(open("gpurun_out/m_bench_cfg$c.json"))
print("cfg$c", d["value"], d["e2e"]["value"], d["update_ms_host_api"], d["solve_ms_host_api"], {k:(v["ms"],v["launches"]) for k,v in d["phases_one_step"].items() if v["launches"] and ("fwd" in k or "bwd" in k)})
PY
done
