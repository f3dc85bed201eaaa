"""dev tool: per-level timeline of one overlapped update! (device-side globaltimer stamps)."""
import os, sys, json
os.environ["TLPB200_TRACE_FACTOR"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
lp = lpgen.config(int(cfg) if cfg.isdigit() else cfg); A = lp.A; m, n = A.shape
kkt = pkg.setup(A, pkg.K1() if cfg != "3" else pkg.K2(), pkg.Backend())
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-3, 3, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
for _ in range(3):
    kkt.update(th, rP, rD)
tr = kkt.factor_trace().astype(np.float64)
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, (tr - t0) / 1e3, np.nan)     # us
names = ["diag", "trsm", "urg", "lazy"]
print("level | " + " | ".join(f"{n:>5s} start   end" for n in names))
for l in range(tr.shape[0]):
    if l % 4 == 0 or l > tr.shape[0] - 6:
        print(f"{l:5d} | " + " | ".join(f"{tr[l, c, 0]:9.1f} {tr[l, c, 1]:8.1f}" for c in range(4)))
d = tr[:, 0, 0]
print("diag-start to diag-start per level (us): first 10", np.round(np.diff(d)[:10], 1), "\nmid", np.round(np.diff(d)[35:45], 1), "\nlast 10", np.round(np.diff(d)[-10:], 1))
print("total span us", np.nanmax(tr))
lz = tr[:, 3, :]
busy = np.nansum(lz[:, 1] - lz[:, 0])
print("sum of lazy spans us", busy)
np.save(os.path.join(ROOT, "gpurun_out", f"factor_trace_cfg{cfg}.npy"), tr)
