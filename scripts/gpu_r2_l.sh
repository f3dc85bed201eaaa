#!/bin/bash
# round-2 GPU session L: final validation at the driver's settings -- smoke(), full GPU test-suite, both bench arms on config T
# (--steps 20 --warmup 5), then every other BASELINE config (CPU arm first)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/l_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/l_ref_cfgT.json 2> gpurun_out/l_ref_cfgT.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/l_bench_cfgT.json 2> gpurun_out/l_bench_cfgT.err
for c in 2 3 4 5; do
  timeout 600 python bench.py --impl reference --config $c --steps 8 --warmup 3 > gpurun_out/l_ref_cfg$c.json 2> gpurun_out/l_ref_cfg$c.err
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/l_bench_cfg$c.json 2> gpurun_out/l_bench_cfg$c.err
done
tail -3 gpurun_out/l_smoke.log; tail -c 400 gpurun_out/l_pytest.log
for c in T 2 3 4 5; do echo "== cfg$c"; cut -c1-220 gpurun_out/l_bench_cfg$c.json; done
tail -4 gpurun_out/l_ref_cfgT.err gpurun_out/l_bench_cfgT.err
