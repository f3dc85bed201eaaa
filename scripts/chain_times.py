"""dev tool: per-block publish times of the dense sweeps (root supernode of cfg2) + per-class profile of one solve."""
import os, sys, json
os.environ["TLPB200_CHAIN_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
lp = lpgen.config(int(cfg) if cfg.isdigit() else cfg); A = lp.A; m, n = A.shape
kkt = pkg.setup(A, pkg.K1() if cfg != "3" else pkg.K2(), pkg.Backend())
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-3, 3, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n); dx = np.zeros(n); dy = np.zeros(m)
kkt.update(th, rP, rD)
for _ in range(5):
    kkt.solve(dx, dy, xp, xd)
f, b = kkt.chain_times()
out = {}
for name, t in (("fwd", f), ("bwd", b[::-1])):
    t = t[t > 0]
    d = np.diff(t)
    out[name] = {"n": int(len(t)), "total_us": float((t[-1] - t[0]) / 1e3), "steps_ns": d.tolist()}
    print(name, "blocks", len(t), "span %.1f us" % ((t[-1] - t[0]) / 1e3), "median step %.0f ns" % np.median(d),
          "first-quarter mean %.0f  last-quarter mean %.0f" % (d[:len(d) // 4].mean(), d[-len(d) // 4:].mean()))
kkt.set_profiling(True)
kkt.update(th, rP, rD)
kkt.solve(dx, dy, xp, xd)
st = kkt.stats()
cls = {k: (round(a, 4), int(c)) for k, a, c in zip(pkg._lib.KERNEL_CLASSES, st["ms_class"], st["n_class"]) if c}
print(cls)
out["classes"] = cls
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"chain_times_cfg{cfg}.json"), "w"))
