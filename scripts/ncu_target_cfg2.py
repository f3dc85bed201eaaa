"""ncu target: config 2, three rounds of update! + solve! (the third is the one to capture), then one pass of the
device-resident HSD loop.  Used with
  ncu --set full --clock-control none --import-source on -k regex:"k_assemble_k1|k_k1_rhs|k_fwd_big|k_bwd_big|k_k1_recover|k_ipm_" -s 10 -c 14 -o gpurun_out/r02_solve python scripts/ncu_target_cfg2.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader
pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen

lp = lpgen.config(2)
A = lp.A
m, n = A.shape
rng = np.random.default_rng(5)
k = pkg.setup(A, pkg.K1(), pkg.Backend())
theta = np.exp(rng.uniform(-2, 2, n)); regP = np.full(n, 1e-6); regD = np.full(m, 1e-6)
dx = np.zeros(n); dy = np.zeros(m)
for r in range(3):
    k.update(theta, regP, regD)
    k.solve(dx, dy, rng.standard_normal(m), rng.standard_normal(n))
d = pkg.DeviceHSD(k, lp.b, lp.c, lp.l, lp.u)
d.iterate()
print("done", k.stats()["nnzL"])
