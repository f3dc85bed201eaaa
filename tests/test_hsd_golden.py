"""The reference's own known answers for the caller of the KKT path (row a12 of SURVEY 8a), restated from
/root/reference/test/IPM/HSD.jl and checked against BOTH the oracle restatement (oracle/hsd_ref.py) and the host mirror
that drives the product (tulip.jl_b200/hsd.py):
  * max_step_length values                       (test/IPM/HSD.jl:33-42; src/IPM/HSD/step.jl:274-306)
  * convergence check at the optimal point       (test/IPM/HSD.jl:44-88;  src/IPM/HSD/HSD.jl:136-196)
  * residual formulas at an arbitrary point      (test/IPM/HSD.jl:96-140; src/IPM/HSD/HSD.jl:77-128)"""
import numpy as np
import pytest

import tlpb200_loader
from oracle import hsd_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import hsd  # noqa: E402

# test/IPM/HSD.jl:46-63: min x1 - x2  s.t. x1 + x2 = 1, x1 - x2 = 0, 0 <= x <= 2
A = np.array([[1.0, 1.0], [1.0, -1.0]])
b = np.array([1.0, 0.0])
c = np.array([1.0, -1.0])
l = np.array([0.0, 0.0])
u = np.array([2.0, 2.0])
SQRT_EPS = float(np.sqrt(np.finfo(float).eps))


def test_max_step_length_single_vector():                      # test/IPM/HSD.jl:33-38
    one, zero = np.ones(1), np.zeros(1)
    for f in (hsd_ref.max_step_length_vec, hsd._max_step):
        assert f(one, one) == np.inf
        assert f(one, -one) == pytest.approx(1.0)
        assert f(zero, -one) == pytest.approx(0.0)
        assert f(zero, one) == np.inf


def test_max_step_length_point():                               # test/IPM/HSD.jl:10-31,40-42
    pt = hsd_ref.Point(2, 2, 1)
    d = hsd_ref.Point(2, 2, 1)
    for q in (pt, d):
        q.x[:] = 1; q.xl[:] = 1; q.xu[:] = 1; q.y[:] = 0; q.zl[:] = 0; q.zu[:] = 0
        q.tau = 1.0; q.kappa = 1.0; q.mu = 1.0
    assert hsd_ref.max_step_length(pt, d) == pytest.approx(1.0)
    # the host mirror's step length on the same point / direction (step.jl:295-306)
    h = hsd.HSD(A, b, c, l, u, kkt=None)
    h.xl[:] = 1; h.xu[:] = 1; h.zl[:] = 0; h.zu[:] = 0; h.tau = 1.0; h.kappa = 1.0
    D = hsd._Dir(2, 2)
    D.x[:] = 1; D.xl[:] = 1; D.xu[:] = 1; D.zl[:] = 0; D.zu[:] = 0; D.tau = 1.0; D.kappa = 1.0
    assert h._alpha(D) == pytest.approx(1.0)


def _set_optimal(get, setv):
    # test/IPM/HSD.jl:67-78: x1 = x2 = 0.5; xl = 0.5; xu = 1.5; tau = 1; y = (0, 1); zl = zu = 0; kappa = 0
    setv("x", [0.5, 0.5]); setv("xl", [0.5, 0.5]); setv("xu", [1.5, 1.5]); setv("y", [0.0, 1.0])
    setv("zl", [0.0, 0.0]); setv("zu", [0.0, 0.0])


def test_convergence_at_the_optimal_point():                    # test/IPM/HSD.jl:80-88
    dat = hsd_ref.IPMData(A, b, True, c, 0.0, l, u)
    o = hsd_ref.HSDRef(dat, kkt=None, params=hsd_ref.IPMOptions(TolerancePFeas=SQRT_EPS, ToleranceDFeas=SQRT_EPS,
                                                               ToleranceRGap=SQRT_EPS, ToleranceIFeas=SQRT_EPS))
    _set_optimal(None, lambda k, v: getattr(o.pt, k).__setitem__(slice(None), v))
    o.pt.tau, o.pt.kappa, o.pt.mu = 1.0, 0.0, 0.0
    o.compute_residuals()
    o.update_solver_status()
    assert o.status == "Trm_Optimal"

    h = hsd.HSD(A, b, c, l, u, kkt=None, params=hsd.IPMOptions(TolerancePFeas=SQRT_EPS, ToleranceDFeas=SQRT_EPS,
                                                               ToleranceRGap=SQRT_EPS, ToleranceIFeas=SQRT_EPS))
    _set_optimal(None, lambda k, v: getattr(h, k).__setitem__(slice(None), v))
    h.tau, h.kappa, h.mu = 1.0, 0.0, 0.0
    h.compute_residuals()
    h.update_solver_status()
    assert h.status == "Trm_Optimal"


def test_residual_formulas():                                   # test/IPM/HSD.jl:96-140
    x = np.array([3.0, 5.0]); xl = np.array([1.0, 8.0]); xu = np.array([2.0, 1.0])
    y = np.array([10.0, -2.0]); zl = np.array([2.0, 1.0]); zu = np.array([5.0, 7.0])
    tau, kappa = 0.5, 0.1
    want = dict(rp=tau * b - A @ x, rl=tau * l - (x - xl), ru=tau * u - (x + xu), rd=tau * c - A.T @ y - zl + zu,
                rg=c @ x - (b @ y + l @ zl - u @ zu) + kappa)
    dat = hsd_ref.IPMData(A, b, True, c, 0.0, l, u)
    o = hsd_ref.HSDRef(dat, kkt=None)
    o.pt.x[:] = x; o.pt.xl[:] = xl; o.pt.xu[:] = xu; o.pt.y[:] = y; o.pt.zl[:] = zl; o.pt.zu[:] = zu
    o.pt.tau, o.pt.kappa, o.pt.mu = tau, kappa, 0.0
    o.compute_residuals()
    h = hsd.HSD(A, b, c, l, u, kkt=None)
    h.x[:] = x; h.xl[:] = xl; h.xu[:] = xu; h.y[:] = y; h.zl[:] = zl; h.zu[:] = zu
    h.tau, h.kappa, h.mu = tau, kappa, 0.0
    h.compute_residuals()
    for obj in (o, h):
        for k, v in want.items():
            np.testing.assert_allclose(getattr(obj, k), v, rtol=1e-14, atol=1e-14)
