#!/bin/bash
# round-2 GPU session U: final state after the chain-kernel rewrite -- micro-benchmark log, GPU test-suite, compute-sanitizer on
# smoke(), then the bench lines of every BASELINE config (CPU arm first on 2-5; config T: GPU arm at the driver's settings)
mkdir -p gpurun_out
timeout 120 scripts/chain_bench > gpurun_out/u_chain_bench.log 2>&1; echo "chain_bench rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/u_pytest.log
tail -2 gpurun_out/u_pytest.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/u_racecheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/u_memcheck_smoke.log
for f in gpurun_out/u_racecheck_smoke.log gpurun_out/u_memcheck_smoke.log; do echo "== $f"; grep -E "SUMMARY|rc=|SMOKE" $f | tail -4; done
for c in 2 3 4 5; do
  timeout 600 python bench.py --impl reference --config $c --steps 8 --warmup 3 > gpurun_out/u_ref_cfg$c.json 2> gpurun_out/u_ref_cfg$c.err
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/u_bench_cfg$c.json 2> gpurun_out/u_bench_cfg$c.err
  echo "== cfg$c rc=$?"; cut -c1-200 gpurun_out/u_bench_cfg$c.json
done
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/u_bench_cfgT.json 2> gpurun_out/u_bench_cfgT.err
echo "== cfgT rc=$?"; cut -c1-300 gpurun_out/u_bench_cfgT.json; tail -4 gpurun_out/u_bench_cfgT.err
