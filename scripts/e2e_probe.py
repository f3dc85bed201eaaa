import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sp
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
lp = lpgen.config(2); A = lp.A; m, n = A.shape
AT = A.T.tocsc()
kkt = pkg.setup(A, pkg.K1(), pkg.Backend())
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-5, 5, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n); dx = np.zeros(n); dy = np.zeros(m)
for _ in range(2): kkt.update(th, rP, rD)
def timed(tag, pre):
    ts = []
    for rep in range(4):
        pre()
        t0 = time.perf_counter(); kkt.update(th, rP, rD); ts.append((time.perf_counter() - t0) * 1e3)
    print("%-34s %s" % (tag, ["%.1f" % t for t in ts]), flush=True)
timed("nothing", lambda: None)
timed("one solve before", lambda: kkt.solve(dx, dy, xp, xd))
timed("4 solves before", lambda: [kkt.solve(dx, dy, xp, xd) for _ in range(4)])
def churn():
    a = [np.zeros(n) + 1.0 for _ in range(40)]; b = sum(x[0] for x in a); return b
timed("numpy alloc churn (160KB arrays)", churn)
timed("scipy spmv", lambda: (A @ xd, AT @ xp))
timed("np.where/div", lambda: np.where(th > 1, 1.0 / th, 0.0))
timed("dot products", lambda: [xd @ xd for _ in range(20)])
timed("solve + churn", lambda: (kkt.solve(dx, dy, xp, xd), churn()))
timed("solve with fresh rhs arrays", lambda: kkt.solve(np.zeros(n), np.zeros(m), xp.copy(), xd.copy()))
th2 = [np.exp(rng.uniform(-5, 5, n)) for _ in range(4)]
ts = []
for rep in range(4):
    kkt.solve(dx, dy, xp, xd)
    t0 = time.perf_counter(); kkt.update(th2[rep], rP, rD); ts.append((time.perf_counter() - t0) * 1e3)
print("solve then update(new theta array)", ["%.1f" % t for t in ts], flush=True)
