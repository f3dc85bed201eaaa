"""dev tool: update!/solve! wall time (graph mode, synchronous host API) under several environment settings.
usage: sweep_env.py [cfg] "K=V,K=V" "K=V" ..."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
args = sys.argv[1:]
cfg = "2"
if args and "=" not in args[0] and args[0] != "-":
    cfg = args.pop(0)
lp = lpgen.config(int(cfg) if cfg.isdigit() else cfg); A = lp.A; m, n = A.shape
sysk = pkg.K2() if cfg == "3" else pkg.K1()
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-3, 3, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n); dx = np.zeros(n); dy = np.zeros(m)
res = []
ref = None
for a in args or ["-"]:
    env = dict(kv.split("=") for kv in a.split(",")) if a != "-" else {}
    for k in list(os.environ):
        if k.startswith("TLPB200_"):
            del os.environ[k]
    os.environ.update(env)
    kkt = pkg.setup(A, sysk, pkg.Backend())
    for _ in range(2):
        kkt.update(th, rP, rD)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); kkt.update(th, rP, rD); t.append(time.perf_counter() - t0)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); kkt.solve(dx, dy, xp, xd); ts.append(time.perf_counter() - t0)
    sol = np.concatenate([dx, dy])
    if ref is None:
        ref = sol.copy()
    r = {"env": env, "update_ms": round(min(t) * 1e3, 3), "solve_ms": round(min(ts) * 1e3, 3),
         "diff_vs_first": float(np.abs(sol - ref).max() / np.abs(ref).max())}
    res.append(r)
    print(r, flush=True)
    del kkt
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"sweep_env_cfg{cfg}.json"), "w"))
