#!/bin/bash
# round-2 GPU session I: long-column kernels -- GPU test-suite, config 5 bench twice (device-resident loop must converge both times)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
timeout 600 python bench.py --impl reference --config 5 --steps 8 --warmup 3 > gpurun_out/i_ref_cfg5.json 2> gpurun_out/i_ref_cfg5.err
for r in 1 2; do
  timeout 600 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/i_bench_cfg5_run$r.json 2> gpurun_out/i_bench_cfg5_run$r.err
done
tail -c 600 gpurun_out/i_pytest.log
for r in 1 2; do python - <<PY
import json
d=json.load(open("gpurun_out/i_bench_cfg5_run$r.json"))
print("run $r", d["value"], d["e2e"], d["ipm"]["status"], d["ipm"]["iters"], d["ipm_device_resident"].get("status"), d["ipm_device_resident"].get("iters"), d["update_ms_host_api"], d["solve_ms_host_api"])
PY
done
