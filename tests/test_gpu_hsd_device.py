"""GPU tests of the device-resident HSD iteration (include/tlpb200.h tlpb200_hsd_*):
same trajectory as the host mirror on the same KKT backend, same answers as the oracle restatement with the oracle KKT,
the reference's own end-to-end answers for its example LPs (test/examples.jl), reference control flow (regularisation
bump, infeasibility certificates)."""
import numpy as np
import pytest

import tlpb200_loader
from golden.lpex import LPEX
from oracle import hsd_ref, kkt_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd, lpgen  # noqa: E402

pytestmark = pytest.mark.gpu
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))
SYSTEMS = {"K1": pkg.K1, "K2": pkg.K2}


@pytest.mark.parametrize("cfg,sysname", [(2, "K1"), (3, "K2"), (4, "K1"), (5, "K1"), ("T", "K1"), (2, "K2")])
def test_device_trajectory_matches_host_mirror_and_oracle(cfg, sysname):
    lp = lpgen.config(cfg, mini=True)
    # (a) host mirror on the same device KKT backend: the two differ only in the order of floating-point reductions
    kh = pkg.setup(lp.A, SYSTEMS[sysname](), pkg.Backend())
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, kh)
    h.optimize()
    kd = pkg.setup(lp.A, SYSTEMS[sysname](), pkg.Backend())
    d = pkg.DeviceHSD(kd, lp.b, lp.c, lp.l, lp.u)
    assert d.optimize() == h.status == "Trm_Optimal"
    assert d.niter == h.niter
    assert d.n_solve == h.n_solve and d.n_update == h.n_update
    dl, hl = d.log, h.log
    assert len(dl) == len(hl)
    for rd, rh in zip(dl[:4], hl[:4]):          # early iterations: tight (later ones amplify rounding through ill-conditioning)
        for a, b in zip(rd[1:3], rh[1:3]):
            assert abs(a - b) <= 1e-9 * (1 + abs(b)), (rd, rh)
        assert abs(rd[6] - rh[6]) <= 1e-9 * (1 + abs(rh[6]))
    for a, b in ((d.primal_objective, h.primal_objective), (d.dual_objective, h.dual_objective)):
        assert abs(a - b) <= 1e-8 * (1 + abs(b))
    # (b) the oracle restatement with the oracle's own KKT (independent floating point end to end)
    dat = hsd_ref.IPMData(lp.A, lp.b, True, lp.c, 0.0, lp.l, lp.u)
    o = kkt_ref.SparseK1(lp.A) if sysname == "K1" else kkt_ref.SparseK2(lp.A)
    ref = hsd_ref.HSDRef(dat, o)
    ref.optimize()
    assert ref.status == "Trm_Optimal" and abs(ref.niter - d.niter) <= 1
    assert abs(d.primal_objective - ref.primal_objective) <= 1e-7 * (1 + abs(ref.primal_objective))
    # the iterate itself: x / tau of both runs
    pt = d.point()
    np.testing.assert_allclose(pt["x"] / pt["tau"], h.x / h.tau, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(h.x / h.tau).max()))


def test_device_iterate_stepwise_equals_host_iterates():
    """iterate() one pass at a time: after k steps the device iterate equals the host mirror's k-th iterate"""
    lp = lpgen.config(2, mini=True)
    kd = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
    d = pkg.DeviceHSD(kd, lp.b, lp.c, lp.l, lp.u)
    kh = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
    snaps = []
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, kh)
    h.optimize(max_iter=3, callback=lambda hh: snaps.append((hh.x.copy(), hh.y.copy(), hh.zl.copy(), hh.tau, hh.kappa)))
    for k in range(3):
        assert d.iterate() == "Trm_Unknown"
        pt = d.point()
        x, y, zl, tau, kappa = snaps[k]
        # two solver handles = two factorisations per step (atomic accumulation order differs): equal to O(kappa u), not bitwise
        np.testing.assert_allclose(pt["x"], x, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(pt["y"], y, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(pt["zl"], zl, rtol=1e-8, atol=1e-10)
        assert abs(pt["tau"] - tau) <= 1e-9 * abs(tau) and abs(pt["kappa"] - kappa) <= 1e-9 * max(abs(kappa), 1e-3)


@pytest.mark.parametrize("name", list(LPEX))
@pytest.mark.parametrize("sysname", ["K1", "K2"])
def test_example_lps_device_resident(name, sysname):
    """test/examples.jl answers (examples/optimal.jl:37-62, freevars.jl, infeasible.jl, unbounded.jl) from the device loop"""
    lp = LPEX[name]
    dat = hsd_ref.standard_form(**{k: v for k, v in lp.items() if k != "expect"})
    kkt = pkg.setup(dat.A, SYSTEMS[sysname](), pkg.Backend())
    d = pkg.DeviceHSD(kkt, dat.b, dat.c, dat.l, dat.u, c0=dat.c0)
    status = d.optimize()
    exp = lp["expect"]
    tol = 100 * SQRT_EPS
    assert status == exp["status"]
    pt = d.point()
    if "obj" in exp:
        assert abs(d.primal_objective - exp["obj"]) <= tol * (1 + abs(exp["obj"]))
    if "x" in exp:
        np.testing.assert_allclose(pt["x"][:lp["nvar"]] / pt["tau"], exp["x"], atol=tol, rtol=tol)
    if "y" in exp:
        np.testing.assert_allclose(pt["y"] / pt["tau"], exp["y"], atol=tol, rtol=tol)


def test_device_loop_iteration_limit_and_reset():
    lp = lpgen.config(2, mini=True)
    k = pkg.setup(lp.A, pkg.K1(), pkg.Backend())
    d = pkg.DeviceHSD(k, lp.b, lp.c, lp.l, lp.u, params=hsd.IPMOptions(IterationsLimit=3))
    assert d.optimize() == "Trm_IterationLimit" and d.niter == 3
    d.params = hsd.IPMOptions()
    assert d.optimize() == "Trm_Optimal"          # optimize() restarts from the reference's start point
    first = d.primal_objective
    d.reset()
    assert d.optimize() == "Trm_Optimal" and abs(d.primal_objective - first) <= 1e-9 * (1 + abs(first))
