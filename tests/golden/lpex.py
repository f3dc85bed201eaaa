"""Golden LPs: the four example problems the reference's own test-suite solves end to end,
restated as data (NOT copied MPS files), with the answers the reference asserts.

Sources (relative to /root/reference):
  lpex_opt      examples/dat/lpex_opt.mps:3-8, answers examples/optimal.jl:37-62
  lpex_freevars examples/dat/lpex_freevars.mps:3-7, answers examples/freevars.jl:35-57
  lpex_inf      examples/dat/lpex_inf.mps:3-8, answers examples/infeasible.jl:37-54
  lpex_ubd      examples/dat/lpex_ubd.mps:3-6, answers examples/unbounded.jl:34-56

Format: general-form LP  min/max obj'x + obj0  s.t. lcon <= A x <= ucon, lvar <= x <= uvar,
A as COO triplets (0-based).  Explicit zero coefficients present in the MPS are kept
(lpex_inf has X1/ROW3 = 0.), as the reference counts them as non-zeros (ipmdata.jl:121-123).
"""
import numpy as np

INF = np.inf

LPEX = {
    "lpex_opt": dict(
        objsense=True, obj=[1.0, 2.0], obj0=0.0, ncon=2, nvar=2,
        rows=[0, 1, 0, 1], cols=[0, 0, 1, 1], vals=[1.0, 1.0, 1.0, -1.0],
        lcon=[1.0, 0.0], ucon=[1.0, 0.0], lvar=[0.0, 0.0], uvar=[1.0, 1.0],
        expect=dict(status="Trm_Optimal", obj=1.5, x=[0.5, 0.5], y=[1.5, -0.5]),
    ),
    "lpex_freevars": dict(
        objsense=True, obj=[1.0, 1.0, 1.0], obj0=0.0, ncon=3, nvar=3,
        rows=[0, 1, 2, 0, 1, 2, 2], cols=[0, 0, 0, 1, 1, 1, 2], vals=[2.0, 1.0, 1.0, 1.0, 2.0, 1.0, 1.0],
        lcon=[2.0, 2.0, 0.0], ucon=[INF, INF, INF], lvar=[-INF] * 3, uvar=[INF] * 3,
        expect=dict(status="Trm_Optimal", obj=0.0),
    ),
    "lpex_inf": dict(
        objsense=True, obj=[1.0, 1.0], obj0=0.0, ncon=3, nvar=2,
        rows=[0, 1, 2, 0, 1, 2], cols=[0, 0, 0, 1, 1, 1], vals=[1.0, 1.0, 0.0, 1.0, -1.0, 1.0],
        lcon=[1.0, 0.0, 1.0], ucon=[1.0, 0.0, 1.0], lvar=[0.0, 0.0], uvar=[INF, INF],
        expect=dict(status="Trm_PrimalInfeasible"),
    ),
    "lpex_ubd": dict(
        objsense=True, obj=[-1.0, -1.0], obj0=0.0, ncon=1, nvar=2,
        rows=[0, 0], cols=[0, 1], vals=[1.0, -1.0],
        lcon=[1.0], ucon=[1.0], lvar=[0.0, 0.0], uvar=[INF, INF],
        expect=dict(status="Trm_DualInfeasible"),
    ),
}

# The one known-answer vector at the KKT boundary (src/KKT/Test/test.jl:26-44,
# matrix from test/KKT/Cholmod/cholmod.jl:2-5): theta=regP=regD=1, xi_p=xi_d=1.
KKT_CONFORMANCE = dict(
    A=np.array([[1.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 1.0]]),
    dx=np.zeros(4), dy=np.ones(2),
)
