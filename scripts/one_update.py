"""dev tool: cfg2 setup + a few update!/solve! calls (target of ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
lp = lpgen.config(int(cfg) if cfg.isdigit() else cfg); A = lp.A; m, n = A.shape
kkt = pkg.setup(A, pkg.K1() if cfg != "3" else pkg.K2(), pkg.Backend(use_graph=False))
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-5, 5, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n); dx = np.zeros(n); dy = np.zeros(m)
for _ in range(int(os.environ.get("NREP", "2"))):
    kkt.update(th, rP, rD)
    kkt.solve(dx, dy, xp, xd)
print("ok", kkt.stats()["launches_update"], kkt.stats()["launches_solve"])
