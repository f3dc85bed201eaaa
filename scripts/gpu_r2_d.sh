#!/bin/bash
# round-2 GPU session D: GPU test-suite after the dense-column rework, config 5 / config 2 / config 3 bench lines (device-resident
# IPM figures included), racecheck re-run, ncu launch list of the first config-T update!.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
for c in 5 2 3 4; do
  timeout 600 python bench.py --impl reference --config $c --steps 8 --warmup 3 > gpurun_out/d_ref_cfg$c.json 2> gpurun_out/d_ref_cfg$c.err
  timeout 600 python bench.py --config $c --steps 8 --warmup 3 > gpurun_out/d_bench_cfg$c.json 2> gpurun_out/d_bench_cfg$c.err
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/d_racecheck.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/d_launches_cfgT.csv \
    python bench.py --steps 1 --warmup 1 --ipm-limit 0 --no-cpu-baseline > gpurun_out/d_ncu_cfgT.log 2>&1; echo "rc=$?" >> gpurun_out/d_ncu_cfgT.log
tail -c 1200 gpurun_out/d_pytest.log
for c in 5 2 3 4; do echo "== cfg$c"; cut -c1-250 gpurun_out/d_bench_cfg$c.json; done
tail -n 4 gpurun_out/d_racecheck.log; wc -l gpurun_out/d_launches_cfgT.csv
