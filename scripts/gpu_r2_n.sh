#!/bin/bash
# round-2 GPU session N: compute-sanitizer on smoke() (now incl. the device-resident loop and the merged sweeps) and on the
# dense-solve-path parity test
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/n_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/n_racecheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kkt.py -q -x -k "dense_solve_path_parity" > gpurun_out/n_memcheck_dense_solve.log 2>&1; echo "rc=$?" >> gpurun_out/n_memcheck_dense_solve.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_kkt.py -q -x -k "dense_solve_path_parity" > gpurun_out/n_racecheck_dense_solve.log 2>&1; echo "rc=$?" >> gpurun_out/n_racecheck_dense_solve.log
for f in gpurun_out/n_*.log; do echo "== $f"; grep -E "SUMMARY|passed|failed|rc=" $f | tail -4; done
