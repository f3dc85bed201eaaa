"""Standard-form problem data feeding ``KKT.setup`` -- host-side mirror of the reference's ``IPMData``
(/root/reference/src/IPM/ipmdata.jl:14-56 the container, :64-173 the conversion from a row/column ``ProblemData``).

SURVEY 8f-4: lets a benchmark or a test start from a general LP

        min / max  obj'x + obj0    s.t.   lcon <= A0 x <= ucon,   lvar <= x <= uvar

the way the real caller does (``Model.optimize!`` -> ``IPMData(pb, MatrixFactory)`` -> ``HSD(dat, kkt_options)`` ->
``KKT.setup(dat.A, ...)``, src/model.jl:127-131), instead of from a matrix that is already in the form
``A x = b, l <= x <= u`` the KKT boundary sees.  Integer / structural work only; no floating-point hot path here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass
class IPMData:
    """ipmdata.jl:14-56: ``A x = b``, ``l <= x <= u`` (minimisation), bound flags for finite bounds."""
    A: sp.csc_matrix
    b: np.ndarray
    objsense: bool          # True = the original problem is a minimisation (c is already flipped otherwise)
    c: np.ndarray
    c0: float
    l: np.ndarray
    u: np.ndarray

    def __post_init__(self):
        self.A = sp.csc_matrix(self.A, dtype=np.float64)
        self.A.sum_duplicates()
        self.A.sort_indices()
        self.nrow, self.ncol = self.A.shape
        self.b = np.asarray(self.b, np.float64); self.c = np.asarray(self.c, np.float64)
        self.l = np.asarray(self.l, np.float64); self.u = np.asarray(self.u, np.float64)
        if self.b.shape != (self.nrow,) or self.c.shape != (self.ncol,) or self.l.shape != (self.ncol,) or self.u.shape != (self.ncol,):
            raise ValueError("IPMData: inconsistent dimensions")
        self.lflag = np.isfinite(self.l)          # ipmdata.jl:44
        self.uflag = np.isfinite(self.u)          # ipmdata.jl:45


def standard_form(A0, lcon, ucon, lvar, uvar, obj, obj0=0.0, objsense=True) -> IPMData:
    """ipmdata.jl:64-173.  One slack column per non-equality row, appended after the structural columns in row order:

    ==================  =========  ==========================  =======
    row                 slack      slack bounds                b_i
    ==================  =========  ==========================  =======
    lb == ub            none       --                          lb
    free                +s         (-inf, +inf)                0
    a'x <= ub           +s         [0, +inf)                   ub
    a'x >= lb           -s         [0, +inf)                   lb
    lb <= a'x <= ub     +s         [0, ub - lb]                ub
    ==================  =========  ==========================  =======

    A maximisation problem is turned into a minimisation by flipping ``obj`` and ``obj0`` (ipmdata.jl:131-135)."""
    A0 = sp.csc_matrix(A0, dtype=np.float64)
    m, n = A0.shape
    lcon = np.asarray(lcon, np.float64); ucon = np.asarray(ucon, np.float64)
    if lcon.shape != (m,) or ucon.shape != (m,):
        raise ValueError("standard_form: row bounds do not match A0")
    eq = lcon == ucon
    free = np.isneginf(lcon) & np.isposinf(ucon)
    le = np.isneginf(lcon) & np.isfinite(ucon)
    ge = np.isfinite(lcon) & np.isposinf(ucon) & ~eq
    rng = np.isfinite(lcon) & np.isfinite(ucon) & ~eq
    bad = ~(eq | free | le | ge | rng)
    if bad.any():
        i = int(np.nonzero(bad)[0][0])
        raise ValueError(f"Invalid bounds for row {i + 1}: [{lcon[i]}, {ucon[i]}]")       # ipmdata.jl:118
    b = np.where(eq, lcon, np.where(free, 0.0, np.where(ge, lcon, ucon)))
    srow = np.nonzero(~eq)[0]                                   # rows that receive a slack, in row order
    sval = np.where(ge[srow], -1.0, 1.0)
    lslack = np.where(free[srow], -np.inf, 0.0)
    uslack = np.where(rng[srow], ucon[srow] - lcon[srow], np.inf)
    S = sp.csc_matrix((sval, (srow, np.arange(len(srow)))), shape=(m, len(srow)))
    A = sp.hstack([A0, S], format="csc")
    c = np.concatenate([np.asarray(obj, np.float64), np.zeros(len(srow))])
    c0 = float(obj0)
    if not objsense:
        c = -c
        c0 = -c0
    l = np.concatenate([np.asarray(lvar, np.float64), lslack])
    u = np.concatenate([np.asarray(uvar, np.float64), uslack])
    return IPMData(A, b, bool(objsense), c, c0, l, u)
