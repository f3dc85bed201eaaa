"""dev tool: update!/solve! wall time (graph mode, synchronous host API) for several chain-SM reservations."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
lp = lpgen.config(2); A = lp.A; m, n = A.shape
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-3, 3, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n); dx = np.zeros(n); dy = np.zeros(m)
res = []
cases = [(16, -1, 0.5)] + [(e, l, s) for e in (4, 8, 12) for l in (16, 32, 48) for s in (0.4, 0.55)] + [(8, -1, 0.5), (24, -1, 0.5)]
if len(sys.argv) > 1:
    cases = [tuple(float(x) if "." in x else int(x) for x in a.split(",")) for a in sys.argv[1:]]
for early, late, sw in cases:
    os.environ["TLPB200_CHAIN_SMS"] = str(early)
    os.environ["TLPB200_CHAIN_SMS_LATE"] = str(late)
    os.environ["TLPB200_CHAIN_SWITCH"] = str(sw)
    kkt = pkg.setup(A, pkg.K1(), pkg.Backend())
    for _ in range(2):
        kkt.update(th, rP, rD)
    t = []
    for _ in range(4):
        t0 = time.perf_counter(); kkt.update(th, rP, rD); t.append(time.perf_counter() - t0)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); kkt.solve(dx, dy, xp, xd); ts.append(time.perf_counter() - t0)
    res.append({"early": early, "late": late, "switch": sw, "update_ms": round(min(t) * 1e3, 3), "solve_ms": round(min(ts) * 1e3, 3)})
    print(res[-1], flush=True)
    del kkt
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_chain_sms.json"), "w"))
