"""CPU tests that pin the oracle against everything the reference's own tests assert for this path
(SURVEY.md 8c): the KKT conformance vector and the four example LPs."""
import numpy as np
import pytest
import scipy.sparse as sp

from golden.lpex import KKT_CONFORMANCE, LPEX
from oracle import hsd_ref, kkt_ref

SQRT_EPS = float(np.sqrt(np.finfo(float).eps))
TOL = 100 * SQRT_EPS          # examples/optimal.jl:11-12


@pytest.mark.parametrize("cls", [kkt_ref.DenseK1, kkt_ref.SparseK1, kkt_ref.SparseK2])
def test_conformance_vector(cls):
    """src/KKT/Test/test.jl:9-46 on the matrix of test/KKT/Cholmod/cholmod.jl:3-6."""
    A = KKT_CONFORMANCE["A"]
    rp, rd, dx, dy = kkt_ref.run_ls_tests(A, cls(sp.csc_matrix(A)))
    assert rp <= SQRT_EPS and rd <= SQRT_EPS
    np.testing.assert_allclose(dx, KKT_CONFORMANCE["dx"], atol=1e-14)
    np.testing.assert_allclose(dy, KKT_CONFORMANCE["dy"], atol=1e-14)


def _std(lp):
    return hsd_ref.standard_form(**{k: v for k, v in lp.items() if k != "expect"})


@pytest.mark.parametrize("name", list(LPEX))
@pytest.mark.parametrize("cls", [kkt_ref.SparseK1, kkt_ref.SparseK2, kkt_ref.DenseK1])
def test_example_lps(name, cls):
    """examples/{optimal,freevars,infeasible,unbounded}.jl through the restated HSD."""
    lp = LPEX[name]
    dat = _std(lp)
    h = hsd_ref.HSDRef(dat, cls(dat.A))
    status = h.optimize()
    exp = lp["expect"]
    assert status == exp["status"]
    if "obj" in exp:
        assert abs(h.primal_objective - exp["obj"]) <= TOL * (1 + abs(exp["obj"]))
    nv = lp["nvar"]
    if "x" in exp:
        np.testing.assert_allclose(h.pt.x[:nv] / h.pt.tau, exp["x"], atol=TOL, rtol=TOL)
    if "y" in exp:
        np.testing.assert_allclose(h.pt.y / h.pt.tau, exp["y"], atol=TOL, rtol=TOL)
    if name == "lpex_freevars":
        x = h.pt.x[:nv] / h.pt.tau                      # examples/freevars.jl:44-56
        assert 2 * x[0] + x[1] >= 2 - TOL and x[0] + 2 * x[1] >= 2 - TOL and x.sum() >= -TOL
        s = (h.pt.zl - h.pt.zu)[:nv] / h.pt.tau
        np.testing.assert_allclose(s, 0, atol=TOL)
    if name == "lpex_inf":
        y = h.pt.y / h.pt.tau                           # examples/infeasible.jl:44-53 (ray, scale free)
        s = (h.pt.zl - h.pt.zu)[:nv] / h.pt.tau
        sc = max(1.0, np.abs(y).max())
        assert (y[0] + y[2]) / sc >= TOL
        assert abs(y[0] + y[1] + s[0]) / sc <= TOL and abs(y[0] - y[1] + y[2] + s[1]) / sc <= TOL
    if name == "lpex_ubd":
        x = h.pt.x[:nv]                                 # examples/unbounded.jl:41-48 (ray)
        sc = max(1.0, np.abs(x).max())
        assert x.min() / sc >= -TOL and abs(x[0] - x[1]) / sc <= TOL and (-x[0] - x[1]) / sc <= -TOL


def test_backends_agree_random():
    """K1 dense / K1 sparse / K2 restatements solve the same system (KKT.jl:70-75)."""
    rng = np.random.default_rng(5)
    m, n = 40, 90
    A = sp.random(m, n, density=0.15, random_state=3, format="csc") + sp.eye(m, n, format="csc")
    theta = np.exp(rng.uniform(-3, 3, n)); regP = np.full(n, 1e-8); regD = np.full(m, 1e-8)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    sols = []
    for cls in (kkt_ref.DenseK1, kkt_ref.SparseK1, kkt_ref.SparseK2):
        k = cls(A); k.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m); k.solve(dx, dy, xi_p, xi_d)
        rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
        assert rp < 1e-8 and rd < 1e-8
        sols.append(np.concatenate([dx, dy]))
    for s in sols[1:]:
        np.testing.assert_allclose(s, sols[0], rtol=1e-8, atol=1e-10)


def test_splu_path_matches_dense():
    """the SuperLU stand-in (used above dense_below) agrees with the dense factorisations"""
    rng = np.random.default_rng(7)
    m, n = 120, 260
    A = sp.random(m, n, density=0.05, random_state=1, format="csc") + sp.eye(m, n, format="csc")
    theta = np.exp(rng.uniform(-4, 4, n)); regP = np.full(n, 1e-7); regD = np.full(m, 1e-7)
    xi_p = rng.standard_normal(m); xi_d = rng.standard_normal(n)
    out = []
    for cls, kw in ((kkt_ref.SparseK1, dict(dense_below=10)), (kkt_ref.SparseK1, dict(dense_below=10**6)),
                    (kkt_ref.SparseK2, dict(dense_below=10)), (kkt_ref.SparseK2, dict(dense_below=10**6))):
        k = cls(A, **kw); k.update(theta, regP, regD)
        dx = np.zeros(n); dy = np.zeros(m); k.solve(dx, dy, xi_p, xi_d)
        out.append(np.concatenate([dx, dy]))
    for s in out[1:]:
        np.testing.assert_allclose(s, out[0], rtol=1e-7, atol=1e-9)


def test_dimension_and_posdef_errors():
    A = sp.csc_matrix(KKT_CONFORMANCE["A"])
    k = kkt_ref.SparseK1(A)
    with pytest.raises(kkt_ref.DimensionMismatch):
        k.update(np.ones(3), np.ones(4), np.ones(2))          # spd.jl:26-28
    with pytest.raises(kkt_ref.PosDefException):
        k.update(np.ones(4), np.ones(4), -10 * np.ones(2))    # spd.jl:47
