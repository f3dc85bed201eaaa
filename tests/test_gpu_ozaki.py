"""GPU tests of the tcgen05 int8 (Ozaki) Schur-update path (tulip.jl_b200/csrc/kernels_ozaki.cu): exactness of the
digit-plane products against an 80-bit reference, and parity of the factorisation / solves with the FP64 DMMA path and
the oracle on the same inputs (reference: cholesky! in /root/reference/src/KKT/Cholmod/spd.jl:46)."""
import ctypes as C

import numpy as np
import pytest

import tlpb200_loader
from oracle import kkt_ref

pkg = tlpb200_loader.load()
from tulip_jl_b200 import lpgen  # noqa: E402

pytestmark = pytest.mark.gpu


def _ozaki(P, C0, ksplit):
    lib = pkg._lib.load()
    R, K = P.shape
    Pf = np.asfortranarray(P)
    Cf = np.asfortranarray(C0.copy())
    ms = (C.c_float * 3)()
    err = C.c_int32(0)
    rc = lib.tlpb200_debug_ozaki(Pf.ctypes.data_as(C.POINTER(C.c_double)), R, K, Cf.ctypes.data_as(C.POINTER(C.c_double)),
                                 ksplit, 0, ms, C.byref(err))
    assert rc == 0 and err.value == 0
    return Cf


@pytest.mark.parametrize("R,K,ksplit,spread", [(128, 32, 0, 0.0), (300, 200, 64, 0.0), (260, 1024, 256, 8.0), (200, 2048, 0, 3.0)])
def test_digit_plane_products_are_fp64_grade(R, K, ksplit, spread):
    """C -= P P' through int8 digit planes == the 80-bit product rounded once per RED, up to the dropped digit pairs
    (~2^-57 of the row scales sqrt(K_ii K_jj): below FP64 rounding of an ordinary GEMM)."""
    rng = np.random.default_rng(R + K)
    P = rng.standard_normal((R, K))
    if spread:
        P *= np.exp(rng.uniform(-spread, spread, (R, 1)))
        P *= np.exp(rng.uniform(-spread / 2, 0, (R, K)))
    C0 = np.zeros((R, R))
    Cg = _ozaki(P, C0, ksplit)
    Pl = P.astype(np.longdouble)
    ref = -(Pl @ Pl.T)
    low = np.tril(np.ones((R, R), bool))
    d = np.abs(Cg.astype(np.longdouble) - ref)[low].astype(np.float64)
    nrm = np.sqrt(np.outer((P * P).sum(1), (P * P).sum(1)))[low]
    nsplit = max(1, -(-K // ksplit)) if ksplit else 1
    # one rounding per RED relative to the value + the dropped digit pairs: rms sqrt(7 K) * 74^2 * 2^-76 of
    # 2^(E_i+E_j) <= 4 sqrt(K_ii K_jj), generous factor 64
    tol = 2.0 ** -52 * nsplit * np.abs(ref)[low].astype(np.float64) + 64.0 * np.sqrt(7.0 * K) * 5476.0 * 2.0 ** -76 * nrm
    assert np.all(d <= tol)
    assert np.array_equal(Cg[~low], C0[~low])          # the strict upper triangle is never written


def _medium_lp():
    return lpgen.random_sparse(1800, 3600, 8, seed=424242, name="oz_medium")


def test_factor_and_solve_parity_with_fp64_path_and_oracle():
    lp = _medium_lp()
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(3)
    koz = pkg.setup(A, pkg.K1(), pkg.Backend(ozaki_ncol=512))
    kfp = pkg.setup(A, pkg.K1(), pkg.Backend(ozaki_ncol=-1))
    assert koz.stats()["oz_tasks"] > 0 and kfp.stats()["oz_tasks"] == 0
    o = kkt_ref.SparseK1(A)
    for spread in (3.0, 9.0):
        theta = np.exp(rng.uniform(-spread, spread, n)); regP = np.full(n, 1e-8); regD = np.full(m, 1e-8)
        xp = rng.standard_normal(m); xd = rng.standard_normal(n)
        sols = []
        for k in (koz, kfp, o):
            k.update(theta, regP, regD)
            dx = np.zeros(n); dy = np.zeros(m)
            k.solve(dx, dy, xp, xd)
            sols.append((dx, dy))
            rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xp, xd)
            scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
            assert rp <= 1e-8 * scale and rd <= 1e-8 * scale
        lo, xo = koz.debug_lx()
        lf, _ = kfp.debug_lx()
        assert np.abs(lo - lf).max() <= 1e-9 * np.abs(lf).max()
        # the two device paths agree at least as well as each agrees with the oracle (ill-conditioned late-IPM theta)
        for a in (0, 1):
            ref = sols[2][a]
            spread_cpu = np.abs(sols[1][a] - ref).max()
            assert np.abs(sols[0][a] - ref).max() <= max(1e-8 * np.abs(ref).max(), 10.0 * spread_cpu)


def test_profiled_update_uses_the_tensor_path():
    lp = _medium_lp()
    A = lp.A
    m, n = A.shape
    k = pkg.setup(A, pkg.K1(), pkg.Backend(ozaki_ncol=512))
    k.set_profiling(True)
    k.update(np.ones(n), np.full(n, 1e-6), np.full(m, 1e-6))
    st = k.stats()
    cls = dict(zip(pkg._lib.KERNEL_CLASSES, st["n_class"]))
    assert cls["oz_update"] > 0 and cls["oz_slice"] > 0
    assert st["flops_update_oz"] > st["flops_update_ext"] * 0.2


def test_breakdown_with_the_tensor_path_raises_and_recovers():
    """spd.jl:47 / step.jl:34-51: a factorisation breakdown inside a supernode that uses the tcgen05 path must surface as
    PosDefException (never a hang or a pipeline time-out), and the retry with larger regularisations must succeed."""
    lp = _medium_lp()
    A = lp.A
    m, n = A.shape
    rng = np.random.default_rng(5)
    k = pkg.setup(A, pkg.K1(), pkg.Backend(ozaki_ncol=512))
    theta = np.exp(rng.uniform(-2, 2, n)); regP = np.full(n, 1e-6)
    with pytest.raises(pkg.PosDefException):
        k.update(theta, regP, np.full(m, -1e6))            # K = A D A' - 1e6 I is indefinite
    regD = np.full(m, 1e-6)
    k.update(theta, regP, regD)
    xp = rng.standard_normal(m); xd = rng.standard_normal(n)
    dx = np.zeros(n); dy = np.zeros(m)
    k.solve(dx, dy, xp, xd)
    rp, rd = kkt_ref.kkt_residuals(A, theta, regP, regD, dx, dy, xp, xd)
    scale = max(1.0, np.abs(dx).max(), np.abs(dy).max())
    assert rp <= 1e-9 * scale and rd <= 1e-9 * scale
