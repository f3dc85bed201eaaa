#!/bin/bash
# round-2 GPU session K: ncu --set full of the HBM-bound kernels (assemble, rhs / recovery, dense sweeps, IPM kernels) on config 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_k1|k_k1_rhs|k_fwd_big|k_bwd_big|k_k1_recover|k_ipm_residuals|k_ipm_theta|k_ipm_newton" -s 10 -c 12 -f -o gpurun_out/r02_solve python scripts/ncu_target_cfg2.py > gpurun_out/k_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/k_ncu.log
tail -5 gpurun_out/k_ncu.log; ls -la gpurun_out/r02_solve.ncu-rep
