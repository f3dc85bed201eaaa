"""dev tool: one update! of the mid-size config (target of the ncu capture of the 128x128 tcgen05 kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
lp = lpgen.random_sparse(30000, 60000, 7, seed=777, name="mid_random_3e4"); A = lp.A; m, n = A.shape
kkt = pkg.setup(A, pkg.K1(), pkg.Backend(use_graph=False))
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-5, 5, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
kkt.update(th, rP, rD)
print("ok", kkt.stats()["launches_update"])
