"""Ad-hoc GPU parity check (dev tool): device KKT path vs the oracle on small/medium configs."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.sparse as sp

import tlpb200_loader

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402
from oracle import kkt_ref  # noqa: E402


def one(lp, sysname, rng, big=False, backend=None):
    A = lp.A
    m, n = A.shape
    sy = pkg.K1() if sysname == "K1" else pkg.K2()
    t0 = time.time()
    k = pkg.setup(A, sy, backend or pkg.Backend())
    t1 = time.time()
    theta = np.exp(rng.uniform(-6, 6, n))
    regP = np.full(n, 1e-6)
    regD = np.full(m, 1e-6)
    xi_p = rng.standard_normal(m)
    xi_d = rng.standard_normal(n)
    k.update(theta, regP, regD)
    t2 = time.time()
    k.update(theta, regP, regD)
    t2b = time.time()
    dx = np.zeros(n)
    dy = np.zeros(m)
    k.solve(dx, dy, xi_p, xi_d)
    t3 = time.time()
    k.solve(dx, dy, xi_p, xi_d)
    t3b = time.time()
    rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
    msg = (f"{lp.name:18s} {sysname} m={m} n={n} setup={t1-t0:.2f}s upd={t2-t1:.3f}/{t2b-t2:.4f}s "
           f"solve={t3-t2b:.3f}/{t3b-t3:.4f}s |rp|={rp:.2e} |rd|={rd:.2e}")
    if not big:
        o = kkt_ref.SparseK1(A) if sysname == "K1" else kkt_ref.SparseK2(A)
        o.update(theta, regP, regD)
        dx0 = np.zeros(n)
        dy0 = np.zeros(m)
        o.solve(dx0, dy0, xi_p, xi_d)
        ex = np.linalg.norm(dx - dx0, np.inf) / max(1e-300, np.linalg.norm(dx0, np.inf))
        ey = np.linalg.norm(dy - dy0, np.inf) / max(1e-300, np.linalg.norm(dy0, np.inf))
        rp0, rd0 = kkt_ref.kkt_residuals(A, theta, regP, regD, dx0, dy0, xi_p, xi_d)
        msg += f" relerr dx={ex:.2e} dy={ey:.2e} (oracle res {rp0:.1e},{rd0:.1e})"
    print(msg, flush=True)
    st = k.stats()
    print("    ", {q: st[q] for q in ("nnzL", "flops", "nsuper", "npieces", "nlevels", "launches_update", "launches_solve")},
          flush=True)
    return k


if __name__ == "__main__":
    from golden.lpex import KKT_CONFORMANCE
    rng = np.random.default_rng(1)
    A = sp.csc_matrix(KKT_CONFORMANCE["A"])
    for sy in (pkg.K1(), pkg.K2()):
        k = pkg.setup(A, sy, pkg.Backend())
        print("conformance", type(sy).__name__, kkt_ref.run_ls_tests(KKT_CONFORMANCE["A"], k), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "mini"
    if which in ("mini", "all"):
        for cfg in (2, 3, 4, 5, "T"):
            lp = lpgen.config(cfg, mini=True)
            for sn in ("K1", "K2"):
                one(lp, sn, rng)
    if which in ("mid", "all"):
        one(lpgen.random_sparse(2000, 4000, 8, name="rand2000"), "K1", rng)
        one(lpgen.random_sparse(2000, 4000, 8, name="rand2000"), "K2", rng)
        one(lpgen.banded_random(20000, 40000, 5, 256, name="banded2e4"), "K1", rng, big=True)
        one(lpgen.staircase(stages=16, nodes=200, arcs=300, name="stair16"), "K2", rng)
    if which in ("big", "all"):
        one(lpgen.config(2), "K1", rng, big=True)
        one(lpgen.config(3), "K2", rng, big=True)
