"""CPU restatement of Tulip's KKT backends (TEST INFRASTRUCTURE, see oracle/__init__.py).

Each class restates one reference backend; line numbers are relative to
/root/reference (Tulip.jl v0.9.8).

* ``DenseK1``   <- src/KKT/Dense/lapack.jl:52-119      (dense normal equations, LAPACK Cholesky)
* ``SparseK1``  <- src/KKT/Cholmod/spd.jl:5-70         (sparse A*D*A' + Rd, Cholesky, solve)
* ``SparseK2``  <- src/KKT/Cholmod/sqd.jl:5-74 and
                   src/KKT/LDLFactorizations/ldlfact.jl:63-139 (augmented system, LDL')
* ``run_ls_tests`` <- src/KKT/Test/test.jl:9-46

The sparse factorisations themselves live in third-party code that is absent from
/root/reference (SuiteSparse CHOLMOD via Julia's SparseArrays stdlib, version = the
running Julia's, not pinned; LDLFactorizations.jl compat "0.10", Project.toml:28).  Their
published algorithm is "fill-reducing symmetric permutation, then Cholesky / LDL' without
numerical pivoting"; the *solution* of the linear system does not depend on which
permutation is used, so the oracle factors with SciPy: dense ``cho_factor`` for small
systems and SuperLU ``splu`` (symmetric mode, diagonal pivoting) for larger ones.  Any timing taken
from this file is a "CPU stand-in (SciPy/SuperLU) -- NOT CHOLMOD".
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class PosDefException(Exception):
    """Mirror of Julia's LinearAlgebra.PosDefException (thrown at spd.jl:47, ldlfact.jl:115)."""


class DimensionMismatch(Exception):
    """Mirror of Julia's DimensionMismatch (spd.jl:26-34)."""


def _check_dims(kkt, theta, regP, regD):
    # spd.jl:26-34 / sqd.jl:28-36 / lapack.jl:69-77
    if len(theta) != kkt.n:
        raise DimensionMismatch(f"length(θ)={len(theta)} but KKT solver has n={kkt.n}.")
    if len(regP) != kkt.n:
        raise DimensionMismatch(f"length(regP)={len(regP)} but KKT solver has n={kkt.n}")
    if len(regD) != kkt.m:
        raise DimensionMismatch(f"length(regD)={len(regD)} but KKT solver has m={kkt.m}")


class DenseK1:
    """lapack.jl:52-119.  K = (A sqrt(D)) (A sqrt(D))' + Rd ; Cholesky ; two triangular solves."""

    system = "K1"

    def __init__(self, A):
        self.A = np.asarray(A.todense() if sp.issparse(A) else A, dtype=np.float64)
        self.m, self.n = self.A.shape
        self.theta = np.ones(self.n)
        self.regP = np.ones(self.n)
        self.regD = np.ones(self.m)
        self.c = None

    def update(self, theta, regP, regD):
        _check_dims(self, theta, regP, regD)
        self.theta[:] = theta
        self.regP[:] = regP
        self.regD[:] = regD
        sqD = np.sqrt(1.0 / (self.theta + self.regP))          # lapack.jl:86
        B = self.A * sqD[None, :]                               # lapack.jl:87
        K = B @ B.T                                             # lapack.jl:88
        K[np.diag_indices(self.m)] += self.regD                 # lapack.jl:90-92
        try:
            self.c = sla.cho_factor(K, lower=True, check_finite=False)   # lapack.jl:95
        except np.linalg.LinAlgError as e:
            raise PosDefException(str(e))
        if not np.all(np.isfinite(self.c[0][np.diag_indices(self.m)])):
            raise PosDefException("non-finite pivot")

    def solve(self, dx, dy, xi_p, xi_d):
        D = 1.0 / (self.theta + self.regP)                      # lapack.jl:104
        xi = xi_p + self.A @ (D * xi_d)                         # lapack.jl:105-106
        dy[:] = sla.cho_solve(self.c, xi, check_finite=False)   # lapack.jl:109-110
        dx[:] = D * (self.A.T @ dy - xi_d)                      # lapack.jl:113-115


class SparseK1:
    """spd.jl:5-70.  S = A*D*A' + spdiagm(regD) ; factor ; dy = F \\ (xi_p + A D xi_d)."""

    system = "K1"

    def __init__(self, A, dense_below=400):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.m, self.n = self.A.shape
        self.theta = np.ones(self.n)
        self.regP = np.ones(self.n)
        self.regD = np.ones(self.m)
        self._dense = self.m <= dense_below
        self.F = None
        self.nnz_factor = None

    def update(self, theta, regP, regD):
        _check_dims(self, theta, regP, regD)
        self.theta[:] = theta                                    # spd.jl:36-38
        self.regP[:] = regP
        self.regD[:] = regD
        D = 1.0 / (self.theta + self.regP)                       # spd.jl:42
        S = (self.A @ sp.diags(D) @ self.A.T + sp.diags(self.regD)).tocsc()   # spd.jl:43
        self.S = S
        if self._dense:
            try:
                self.F = ("chol", sla.cho_factor(S.toarray(), lower=True, check_finite=False))
            except np.linalg.LinAlgError as e:                   # spd.jl:47
                raise PosDefException(str(e))
        else:
            # stand-in for cholesky!(F, Symmetric(K)) (spd.jl:46): SuperLU, symmetric mode,
            # diagonal pivots (diag_pivot_thresh=0) on an SPD matrix == LDL' up to rounding.
            lu = spla.splu(S, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
            d = lu.U.diagonal()
            if not np.all(np.isfinite(d)) or np.any(d <= 0.0):   # spd.jl:47
                raise PosDefException("non-positive pivot")
            self.F = ("lu", lu)
            self.nnz_factor = lu.L.nnz

    def _fsolve(self, rhs):
        kind, f = self.F
        return sla.cho_solve(f, rhs, check_finite=False) if kind == "chol" else f.solve(rhs)

    def solve(self, dx, dy, xi_p, xi_d):
        D = 1.0 / (self.theta + self.regP)                       # spd.jl:55
        xi = xi_p + self.A @ (D * xi_d)                          # spd.jl:56-57
        dy[:] = self._fsolve(xi)                                 # spd.jl:61
        dx[:] = D * (self.A.T @ dy - xi_d)                       # spd.jl:64-66


class SparseK2:
    """sqd.jl:5-74 (and ldlfact.jl:63-139).  K = [-(theta+Rp) A'; A Rd], rhs [xi_d; xi_p]."""

    system = "K2"

    def __init__(self, A, dense_below=400):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.m, self.n = self.A.shape
        self.theta = np.ones(self.n)
        self.regP = np.ones(self.n)
        self.regD = np.ones(self.m)
        self._dense = (self.m + self.n) <= dense_below
        self.F = None

    def update(self, theta, regP, regD):
        _check_dims(self, theta, regP, regD)
        self.theta[:] = theta
        self.regP[:] = regP
        self.regD[:] = regD
        n, m = self.n, self.m
        # sqd.jl:44-51: diagonal of the (1,1) block is -theta-regP, of the (2,2) block regD
        K = sp.bmat([[sp.diags(-(self.theta + self.regP)), self.A.T],
                     [self.A, sp.diags(self.regD)]], format="csc")
        self.K = K
        if self._dense:
            Kd = K.toarray()
            # LDL' without pivoting of a quasi-definite matrix (sqd.jl:53 / ldlfact.jl:113)
            L, d = _dense_ldl_nopiv(Kd)
            # expected inertia: first n pivots < 0, last m > 0 (ldlfact.jl:112-117 SQDException)
            if (not np.all(np.isfinite(d))) or np.any(d[:n] >= 0) or np.any(d[n:] <= 0):
                raise PosDefException("wrong pivot sign")
            self.F = ("ldl", (L, d))
        else:
            lu = spla.splu(K, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
            d = lu.U.diagonal()
            if not np.all(np.isfinite(d)) or np.any(d == 0.0):
                raise PosDefException("zero pivot")
            self.F = ("lu", lu)

    def solve(self, dx, dy, xi_p, xi_d):
        n = self.n
        xi = np.concatenate([xi_d, xi_p])                        # sqd.jl:61-62
        kind, f = self.F
        if kind == "ldl":
            L, d = f
            z = sla.solve_triangular(L, xi, lower=True, unit_diagonal=True, check_finite=False)
            z /= d
            delta = sla.solve_triangular(L.T, z, lower=False, unit_diagonal=True, check_finite=False)
        else:
            delta = f.solve(xi)                                  # sqd.jl:66
        dx[:] = delta[:n]                                        # sqd.jl:69-70
        dy[:] = delta[n:]


def _dense_ldl_nopiv(K):
    """Textbook right-looking LDL' without pivoting (what ldl_factorize! computes, dense)."""
    N = K.shape[0]
    L = np.array(K, dtype=np.float64)
    d = np.zeros(N)
    for j in range(N):
        d[j] = L[j, j]
        if d[j] == 0.0 or not np.isfinite(d[j]):
            d[j:] = np.nan
            break
        L[j + 1:, j] /= d[j]
        if j + 1 < N:
            L[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j]) * d[j]
    L = np.tril(L, -1) + np.eye(N)
    return L, d


def kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d):
    """test.jl:39-40: rp = A dx + Rd dy - xi_p ; rd = -(theta+Rp) dx + A' dy - xi_d."""
    rp = A @ dx + regD * dy - xi_p
    rd = -dx * (theta + regP) + A.T @ dy - xi_d
    return np.linalg.norm(rp, np.inf), np.linalg.norm(rd, np.inf)


def run_ls_tests(A, kkt, atol=np.sqrt(np.finfo(np.float64).eps)):
    """src/KKT/Test/test.jl:9-46 -- the conformance test every backend must pass.

    ``kkt`` may be any object with ``update(theta, regP, regD)`` and
    ``solve(dx, dy, xi_p, xi_d)``.  Returns (|rp|_inf, |rd|_inf, dx, dy).
    """
    assert hasattr(kkt, "update") and hasattr(kkt, "solve")      # test.jl:19-20
    m, n = A.shape
    theta = np.ones(n)                                           # test.jl:26-28
    regP = np.ones(n)
    regD = np.ones(m)
    kkt.update(theta, regP, regD)                                # test.jl:29
    xi_p = np.ones(m)                                            # test.jl:32-33
    xi_d = np.ones(n)
    dx = np.zeros(n)
    dy = np.zeros(m)
    kkt.solve(dx, dy, xi_p, xi_d)                                # test.jl:36
    rp, rd = kkt_residuals(A, theta, regP, regD, dx, dy, xi_p, xi_d)
    assert rp <= atol, f"|rp|={rp}"                              # test.jl:42
    assert rd <= atol, f"|rd|={rd}"                              # test.jl:43
    return rp, rd, dx, dy
