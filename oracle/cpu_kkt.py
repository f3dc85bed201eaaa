"""CPU supernodal KKT backend (TEST INFRASTRUCTURE + the bench's CPU baseline / ``--impl reference`` arm).

``CpuSupernodalKKT`` follows the reference's CholmodSolver line by line on the Julia side
(src/KKT/Cholmod/spd.jl:22-70 for K1, sqd.jl:24-74 for K2) and replaces the un-vendored CHOLMOD
calls by oracle/cpu_supernodal.c (left-looking supernodal factorisation on OpenBLAS).  The
fill-reducing permutation / supernode partition are taken from the product's host symbolic
analysis (``analyze_only`` -- integer work that tests/test_symbolic.py checks bit-exactly against
brute force); all floating-point work here is independent of the product.

Label for every number produced here: "CPU port: own supernodal Cholesky + SciPy-OpenBLAS, N threads
-- NOT Tulip/CHOLMOD".
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np
import scipy
import scipy.sparse as sp

from .kkt_ref import DimensionMismatch, PosDefException

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load(nthreads):
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libcpu_supernodal.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle`")
        lib = C.CDLL(path)
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
        if not cands:
            raise ImportError("SciPy's bundled OpenBLAS not found")
        rc = lib.cpu_sn_init(os.path.abspath(cands[0]).encode(), int(nthreads))
        if rc != 0:
            raise ImportError(f"cpu_sn_init failed ({rc})")
        _lib = lib
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


class CpuSupernodalKKT:
    def __init__(self, A, system="K1", nthreads=None, symbolic_from=None):
        """``symbolic_from``: a product solver created with Backend(analyze_only=True) for the
        same A/system (only its integer analysis is read)."""
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.AT = self.A.T.tocsc()
        self.m, self.n = self.A.shape
        self.system = system
        self.nthreads = nthreads or os.cpu_count() or 1
        self.lib = _load(self.nthreads)
        sym = symbolic_from.symbolic()
        self.perm = sym["perm"].astype(np.int64)
        self.N = len(self.perm)
        self.iperm = np.empty(self.N, np.int64); self.iperm[self.perm] = np.arange(self.N)
        self.sn_first = np.ascontiguousarray(sym["sn_first"], np.int32)
        self.sn_rowptr = np.ascontiguousarray(sym["sn_rowptr"], np.int64)
        self.sn_rows = np.ascontiguousarray(sym["sn_rows"], np.int32)
        self.nsuper = len(self.sn_first) - 1
        ncol = np.diff(self.sn_first).astype(np.int64)
        nrow = np.diff(self.sn_rowptr)
        self.sn_xptr = np.concatenate([[0], np.cumsum(ncol * nrow)]).astype(np.int64)
        self.col2sn = np.repeat(np.arange(self.nsuper, dtype=np.int32), ncol)
        if system == "K1":
            self.sign = np.ones(self.N, np.int8)
        else:
            self.sign = np.where(self.perm < self.n, -1, 1).astype(np.int8)
        self.Lx = np.zeros(int(self.sn_xptr[-1]))
        self.theta = np.ones(self.n); self.regP = np.ones(self.n); self.regD = np.ones(self.m)

    def update(self, theta, regP, regD):
        if len(theta) != self.n or len(regP) != self.n or len(regD) != self.m:    # spd.jl:26-34
            raise DimensionMismatch("update!: wrong vector length")
        self.theta[:] = theta; self.regP[:] = regP; self.regD[:] = regD           # spd.jl:36-38
        if self.system == "K1":
            D = 1.0 / (self.theta + self.regP)                                    # spd.jl:42
            K = (self.A @ sp.diags(D) @ self.AT + sp.diags(self.regD)).tocsc()    # spd.jl:43
        else:                                                                     # sqd.jl:44-51
            K = sp.bmat([[sp.diags(-(self.theta + self.regP)), self.AT], [self.A, sp.diags(self.regD)]], format="csc")
        Kp = K[self.perm][:, self.perm].tocsc()          # what CHOLMOD does internally with its permutation
        cp = np.ascontiguousarray(Kp.indptr, np.int64); ri = np.ascontiguousarray(Kp.indices, np.int32)
        rc = self.lib.cpu_sn_scatter(C.c_int32(self.N), C.c_int32(self.nsuper), _p(self.sn_first), _p(self.sn_rowptr),
                                     _p(self.sn_rows), _p(self.sn_xptr), _p(cp), _p(ri), _p(Kp.data), _p(self.Lx))
        if rc:
            raise MemoryError("cpu_sn_scatter")
        bad = C.c_int32(-1)
        rc = self.lib.cpu_sn_factor(C.c_int32(self.N), C.c_int32(self.nsuper), _p(self.sn_first), _p(self.sn_rowptr),
                                    _p(self.sn_rows), _p(self.sn_xptr), _p(self.col2sn), _p(self.sign), _p(self.Lx),
                                    C.byref(bad))                                 # spd.jl:46 / sqd.jl:53
        if rc == 1:
            raise PosDefException(f"bad pivot at permuted column {bad.value}")   # spd.jl:47
        if rc:
            raise MemoryError("cpu_sn_factor")

    def _fsolve(self, rhs, mode=0, permuted_in=False, permuted_out=False):
        """(L S L')^{-1} rhs; mode 1 = L^{-1} only, 2 = L^{-T} S only (vectors then live in permuted numbering)"""
        x = np.ascontiguousarray(rhs if permuted_in else rhs[self.perm], dtype=np.float64).copy()
        rc = self.lib.cpu_sn_solve_part(C.c_int32(self.N), C.c_int32(self.nsuper), _p(self.sn_first), _p(self.sn_rowptr),
                                        _p(self.sn_rows), _p(self.sn_xptr), _p(self.sign), _p(self.Lx), _p(x), C.c_int(mode))
        if rc:
            raise MemoryError("cpu_sn_solve")
        if permuted_out:
            return x
        out = np.empty_like(x); out[self.perm] = x
        return out

    def solve(self, dx, dy, xi_p, xi_d):
        if self.system == "K1":
            D = 1.0 / (self.theta + self.regP)                                    # spd.jl:55
            xi = xi_p + self.A @ (D * xi_d)                                       # spd.jl:56-57
            dy[:] = self._fsolve(xi)                                              # spd.jl:61
            dx[:] = D * (self.AT @ dy - xi_d)                                     # spd.jl:64-66
        else:
            d = self._fsolve(np.concatenate([xi_d, xi_p]))                        # sqd.jl:61-66
            dx[:] = d[:self.n]; dy[:] = d[self.n:]                                # sqd.jl:69-70
