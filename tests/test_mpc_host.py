"""CPU tests of the second client of the KKT boundary and of the standard-form build:

* tulip.jl_b200/ipmdata.standard_form == the oracle's restatement of ipmdata.jl:64-173 on the reference's example LPs and
  on random general-form LPs (every row kind: equality, free, <=, >=, ranged; max -> min flip);
* tulip.jl_b200/mpc.MPC holds the reference's own known answers for MPC (test/IPM/MPC.jl: residual formulas :96-140,
  convergence at the optimal point :44-88) and the end-to-end answers test/examples.jl asserts for
  ``IPM_Factory = Factory(MPC)`` (lpex_opt: obj 3/2, x = (1/2, 1/2), y = (3/2, -1/2); lpex_freevars: obj 0), driven here
  with the oracle's KKT backends (dense LAPACK / SuperLU) -- the GPU suite repeats them with the B200 backend;
* call pattern: 1 update! + 2 solve! for the starting point (MPC.jl:359-363), then 1 update! + (2 .. 2 + CorrectionLimit)
  solve! per iteration."""
import numpy as np
import pytest
import scipy.sparse as sp

import tlpb200_loader
from golden.lpex import LPEX
from oracle import hsd_ref, kkt_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd, ipmdata, mpc  # noqa: E402

SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


def _general(lp):
    A0 = sp.coo_matrix((lp["vals"], (lp["rows"], lp["cols"])), shape=(lp["ncon"], lp["nvar"]))
    return ipmdata.standard_form(A0, lp["lcon"], lp["ucon"], lp["lvar"], lp["uvar"], lp["obj"], lp["obj0"], lp["objsense"])


@pytest.mark.parametrize("name", list(LPEX))
def test_standard_form_matches_oracle_on_examples(name):
    lp = LPEX[name]
    mine = _general(lp)
    ref = hsd_ref.standard_form(**{k: v for k, v in lp.items() if k != "expect"})
    assert (mine.A != sp.csc_matrix(ref.A)).nnz == 0
    for a, b in ((mine.b, ref.b), (mine.c, ref.c), (mine.l, ref.l), (mine.u, ref.u)):
        np.testing.assert_array_equal(a, b)
    assert mine.c0 == ref.c0 and mine.objsense == ref.objsense
    np.testing.assert_array_equal(mine.lflag, ref.lflag); np.testing.assert_array_equal(mine.uflag, ref.uflag)


def test_standard_form_row_kinds_and_max_flip():
    rng = np.random.default_rng(5)
    m, n = 12, 7
    A0 = sp.random(m, n, density=0.4, random_state=7, format="coo")
    lcon = np.array([1.0, -np.inf, -np.inf, 2.0, -1.0, 0.0] * 2)
    ucon = np.array([1.0, np.inf, 3.0, np.inf, 4.0, 0.0] * 2)
    lvar = rng.uniform(-1, 0, n); uvar = np.where(rng.random(n) < 0.5, np.inf, 2.0)
    obj = rng.standard_normal(n)
    mine = ipmdata.standard_form(A0, lcon, ucon, lvar, uvar, obj, 0.5, objsense=False)
    ref = hsd_ref.standard_form(obj, 0.5, False, A0.row, A0.col, A0.data, m, n, lcon, ucon, lvar, uvar)
    assert (mine.A != sp.csc_matrix(ref.A)).nnz == 0
    for a, b in ((mine.b, ref.b), (mine.c, ref.c), (mine.l, ref.l), (mine.u, ref.u)):
        np.testing.assert_array_equal(a, b)
    assert mine.c0 == ref.c0 == -0.5 and mine.objsense is False
    assert mine.ncol == n + 8                                   # 4 equality rows of 12 get no slack
    with pytest.raises(ValueError):
        ipmdata.standard_form(A0, np.full(m, np.inf), np.full(m, -np.inf), lvar, uvar, obj)   # ipmdata.jl:118


# test/IPM/MPC.jl:46-63: min x1 - x2  s.t. x1 + x2 = 1, x1 - x2 = 0, 0 <= x <= 2
A = np.array([[1.0, 1.0], [1.0, -1.0]]); b = np.array([1.0, 0.0]); c = np.array([1.0, -1.0])
l = np.array([0.0, 0.0]); u = np.array([2.0, 2.0])


def test_mpc_convergence_at_optimal_point():                    # test/IPM/MPC.jl:65-88
    h = mpc.MPC(A, b, c, l, u, kkt=None)
    h.x[:] = [0.5, 0.5]; h.xl[:] = [0.5, 0.5]; h.xu[:] = [1.5, 1.5]; h.y[:] = [0.0, 1.0]; h.zl[:] = 0; h.zu[:] = 0
    h.tau = 1.0; h.kappa = 0.0; h.mu = 0.0
    h.compute_residuals()
    h.update_solver_status()
    assert h.status == "Trm_Optimal"


def test_mpc_residual_formulas():                               # test/IPM/MPC.jl:96-140
    h = mpc.MPC(A, b, c, l, u, kkt=None)
    x = np.array([3.0, 5.0]); xl = np.array([1.0, 8.0]); xu = np.array([2.0, 1.0])
    y = np.array([10.0, -2.0]); zl = np.array([2.0, 1.0]); zu = np.array([5.0, 7.0])
    h.x[:] = x; h.xl[:] = xl; h.xu[:] = xu; h.y[:] = y; h.zl[:] = zl; h.zu[:] = zu
    h.compute_residuals()
    np.testing.assert_allclose(h.rp, b - A @ x)
    np.testing.assert_allclose(h.rl, l - (x - xl))
    np.testing.assert_allclose(h.ru, u - (x + xu))
    np.testing.assert_allclose(h.rd, c - A.T @ y - zl + zu)
    assert h.rp_nrm == np.abs(h.rp).max() and h.rd_nrm == np.abs(h.rd).max()


class _Counting:
    def __init__(self, inner):
        self.inner = inner
        self.calls = []

    def update(self, *a):
        self.calls.append("u")
        self.inner.update(*a)

    def solve(self, *a):
        self.calls.append("s")
        self.inner.solve(*a)


@pytest.mark.parametrize("name", ["lpex_opt", "lpex_freevars"])
@pytest.mark.parametrize("oracle_cls", [kkt_ref.SparseK1, kkt_ref.SparseK2])
def test_mpc_example_lps_with_oracle_kkt(name, oracle_cls):
    """test/examples.jl:5-19 with IPM_Factory = Factory(MPC): the answers examples/optimal.jl:37-62 / freevars.jl:35-57 assert"""
    lp = LPEX[name]
    dat = _general(lp)
    k = _Counting(oracle_cls(dat.A))
    h = mpc.MPC(dat.A, dat.b, dat.c, dat.l, dat.u, k, c0=dat.c0, objsense=dat.objsense)
    assert h.optimize() == "Trm_Optimal"
    exp = lp["expect"]
    tol = 100 * SQRT_EPS
    assert abs(h.primal_objective - exp["obj"]) <= tol * (1 + abs(exp["obj"]))
    if "x" in exp:
        np.testing.assert_allclose(h.x[:lp["nvar"]], exp["x"], atol=tol, rtol=tol)
    if "y" in exp:
        np.testing.assert_allclose(h.y, exp["y"], atol=tol, rtol=tol)
    # call pattern of the second client (MPC.jl:359-363, MPC/step.jl:58,64,85-109)
    calls = "".join(k.calls)
    assert calls.startswith("uss")
    body = calls[3:].split("u")[1:]
    assert len(body) == h.niter and all(2 <= len(seg) <= 2 + h.params.CorrectionLimit + 1 for seg in body)


def test_mpc_and_hsd_agree_on_a_synthetic_lp():
    """two different IPMs, one LP: same optimal objective (both through the same oracle KKT backend)"""
    from tulip_jl_b200 import lpgen
    lp = lpgen.config(3, mini=True)                  # finite upper bounds: both theta_l and theta_u paths
    k1 = kkt_ref.SparseK2(lp.A)
    h = hsd.HSD(lp.A, lp.b, lp.c, lp.l, lp.u, k1)
    assert h.optimize() == "Trm_Optimal"
    k2 = kkt_ref.SparseK2(lp.A)
    q = mpc.MPC(lp.A, lp.b, lp.c, lp.l, lp.u, k2)
    assert q.optimize() == "Trm_Optimal"
    assert abs(q.primal_objective - h.primal_objective) <= 1e-6 * (1 + abs(h.primal_objective))
