"""dev tool: update! time of a config under schedule variants (env toggles are read at setup)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
cfg = os.environ.get("CFG", "M")
lp = lpgen.random_sparse(30000, 60000, 7, seed=777, name="mid_random_3e4") if cfg == "M" else lpgen.config(int(cfg))
A = lp.A; m, n = A.shape
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-5, 5, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
k = pkg.setup(A, pkg.K1(), pkg.Backend(use_graph=os.environ.get("NOGRAPH") is None))
ts = []
for _ in range(int(os.environ.get("NUPD", "3"))):
    t0 = time.perf_counter(); k.update(th, rP, rD); ts.append(time.perf_counter() - t0)
print(os.environ.get("TAG", ""), [round(t * 1e3, 2) for t in ts], flush=True)
