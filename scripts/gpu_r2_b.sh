#!/bin/bash
# round-2 GPU session B: full GPU test-suite (incl. the device-resident HSD loop), config-5 bench, config T both arms,
# compute-sanitizer memcheck / racecheck of smoke().
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 600 python bench.py --impl reference --config 5 --steps 8 --warmup 3 > gpurun_out/b_ref_cfg5.json 2> gpurun_out/b_ref_cfg5.err
timeout 600 python bench.py --config 5 --steps 8 --warmup 3 > gpurun_out/b_bench_cfg5.json 2> gpurun_out/b_bench_cfg5.err
( time timeout 1200 python bench.py --impl reference --steps 4 --warmup 3 ) > gpurun_out/b_ref_cfgT.json 2> gpurun_out/b_ref_cfgT.err
( time timeout 900 python bench.py --steps 4 --warmup 3 ) > gpurun_out/b_bench_cfgT.json 2> gpurun_out/b_bench_cfgT.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/b_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/b_racecheck.log
tail -c 1500 gpurun_out/b_pytest.log
for f in gpurun_out/b_bench_cfg5.json gpurun_out/b_ref_cfgT.json gpurun_out/b_bench_cfgT.json; do echo "== $f"; cut -c1-300 $f; done
tail -3 gpurun_out/b_memcheck.log gpurun_out/b_racecheck.log
