"""dev tool: cfg2 factorisation with the tcgen05 int8 path on / off: agreement of L, residuals, update! time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import tlpb200_loader; pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen
from oracle import kkt_ref
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
if cfg == "M":
    lp = lpgen.random_sparse(30000, 60000, 7, seed=777, name="mid_random_3e4")
else:
    lp = lpgen.config(int(cfg) if cfg.isdigit() else cfg)
A = lp.A; m, n = A.shape
rng = np.random.default_rng(0)
th = np.exp(rng.uniform(-5, 5, n)); rP = np.full(n, 1e-6); rD = np.full(m, 1e-6)
xp = rng.standard_normal(m); xd = rng.standard_normal(n)
lx = {}
for name, nc in (("fp64", -1), ("ozaki", 0)):
    k = pkg.setup(A, pkg.K1(), pkg.Backend(ozaki_ncol=nc))
    dx = np.zeros(n); dy = np.zeros(m)
    ts = []
    for _ in range(int(os.environ.get('NUPD', '6'))):
        t0 = time.perf_counter(); k.update(th, rP, rD); ts.append(time.perf_counter() - t0)
    k.solve(dx, dy, xp, xd)
    rp, rd = kkt_ref.kkt_residuals(A, th, rP, rD, dx, dy, xp, xd)
    st = k.stats()
    print(f"{name}: nnzL={st['nnzL']:.3e} flops={st['flops']:.3e} nsuper={st['nsuper']} levels={st['nlevels']} update ms {[round(t * 1e3, 2) for t in ts]} residuals {rp:.2e} {rd:.2e} oz_tasks={st['oz_tasks']} "
          f"oz GB={st['oz_bytes'] / 1e9:.2f} launches={st['launches_update']}", flush=True)
    lx[name] = k.debug_lx()[0]
    if name == "ozaki":
        k.set_profiling(True); k.update(th, rP, rD); sp = k.stats(); k.set_profiling(False)
        cls = dict(zip(pkg._lib.KERNEL_CLASSES, zip(sp["ms_class"], sp["n_class"])))
        print({a: (round(b[0], 3), b[1]) for a, b in cls.items() if b[1]}, flush=True)
        print(f"oz_update: {sp['flops_update_oz'] / cls['oz_update'][0] / 1e9:.1f} TF FP64-equivalent (serial, profiling mode); "
              f"DMMA: {sp['flops_update_ext'] / (cls['update'][0] + cls['update128'][0]) / 1e9:.1f} TF", flush=True)
    del k
d = np.abs(lx["fp64"] - lx["ozaki"])
print(f"max |L_fp64 - L_ozaki| = {d.max():.3e} (max |L| = {np.abs(lx['fp64']).max():.3e}); rel Frobenius {np.linalg.norm(d) / np.linalg.norm(lx['fp64']):.3e}")
