"""Host-side mirror of Tulip's KKT plug-in interface for the B200 backend.

The reference boundary is Julia multiple dispatch (/root/reference/src/KKT/KKT.jl):
``setup(A, system, backend)`` (:59), ``update!(kkt, θinv, regP, regD)`` (:83),
``solve!(dx, dy, kkt, ξp, ξd)`` (:100), ``arithmetic`` (:107), ``backend`` (:114),
``linear_system`` (:121).  The same names, argument order and error behaviour are kept here
(``!`` -> trailing underscore), so the parity tests read like test/KKT/*.jl.  All numeric work
happens in libtlpb200.so (CUDA, sm_100a); there is no CPU path in this module.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import _lib


# ---- system / backend tags (src/KKT/systems.jl:6,32,54 ; src/KKT/KKT.jl:16,25) ----------------
class AbstractKKTSystem:
    pass


class DefaultKKTSystem(AbstractKKTSystem):
    """Currently equivalent to K2 (systems.jl:1-6, KKT.jl:134-141)."""


class K1(AbstractKKTSystem):
    """Normal equations (systems.jl:34-54)."""


class K2(AbstractKKTSystem):
    """Augmented system (systems.jl:8-32)."""


@dataclass
class Backend:
    """``TlpB200.Backend <: AbstractKKTBackend``; options live as fields, like
    TlpKrylov.Backend (src/KKT/Krylov/krylov.jl:41-44)."""
    device: int = 0
    ordering: int = 1          # 0 natural, 1 approximate minimum degree
    piece_width: int = 128
    small_elems: int = 4096
    relax_always: int = 8
    use_graph: bool = True
    analyze_only: bool = False  # host symbolic analysis only (CPU tests); numeric calls then fail
    rank: int = 0               # multi-GPU subtree sharding (tulip.jl_b200/parallel.py sets these)
    nranks: int = 1
    dense_col_threshold: int = 0  # K1 dense-column Schur path: 0 auto (max(32, 5% of m)), < 0 off
    dense_solve_ncol: int = 0     # supernodes with >= this many columns use the dense-solve path (0 = default 384)
    ozaki_ncol: int = 0           # K1: supernodes with >= this many columns use the tcgen05 int8 update path (0 = default 1024, < 0 off)
    refine_steps: int = 0         # iterative-refinement steps inside solve!; 0 = the reference's single solve


# ---- exceptions (what the reference throws at this boundary) ---------------------------------
class PosDefException(Exception):
    """LinearAlgebra.PosDefException -- triggers the regularisation bump (HSD/step.jl:40)."""


class DimensionMismatch(Exception):
    """spd.jl:26-34."""


class OutOfMemoryError(MemoryError):
    """mapped to Trm_MemoryLimit by the caller (HSD.jl:327)."""


class TlpB200Error(RuntimeError):
    """CUDA / internal errors: propagate and abort the solve (HSD.jl:333-335)."""


def _raise(code, handle):
    msg = _lib.load().tlpb200_last_error(handle)
    msg = msg.decode() if msg else ""
    if code == _lib.NOT_POSDEF:
        raise PosDefException(msg)
    if code == _lib.OOM:
        raise OutOfMemoryError(msg)
    if code == _lib.BAD_ARG:
        raise DimensionMismatch(msg)
    raise TlpB200Error(f"tlpb200 error {code}: {msg}")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class B200KKTSolver:
    """``B200KKTSolver{Float64,S} <: AbstractKKTSolver{Float64}`` (cf. cholmod.jl:46-60)."""

    def __init__(self, A, system, backend: Backend):
        lib = _lib.load()
        if isinstance(system, DefaultKKTSystem):
            system = K2()                                   # KKT.jl:134-141
        if not isinstance(system, (K1, K2)):
            raise TypeError("system must be K1() or K2()")
        A = sp.csc_matrix(A, dtype=np.float64)              # cholmod.jl:65 convert(SparseMatrixCSC, A)
        A.sum_duplicates()
        A.sort_indices()
        self.A = A                                          # borrowed for the solver's lifetime (spd.jl:19)
        self.m, self.n = A.shape
        self.system = system
        self._sys = _lib.K1 if isinstance(system, K1) else _lib.K2
        self.backend_options = backend
        opt = _lib.Options()
        lib.tlpb200_default_options(C.byref(opt))
        opt.device = backend.device
        opt.ordering = backend.ordering
        opt.piece_width = backend.piece_width
        opt.small_elems = backend.small_elems
        opt.relax_always = backend.relax_always
        opt.use_graph = 1 if backend.use_graph else 0
        opt.analyze_only = 1 if backend.analyze_only else 0
        opt.rank = backend.rank
        opt.nranks = backend.nranks
        opt.dense_col_threshold = backend.dense_col_threshold
        opt.dense_solve_ncol = backend.dense_solve_ncol
        opt.ozaki_ncol = backend.ozaki_ncol
        opt.refine_steps = backend.refine_steps
        colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        rowval = np.ascontiguousarray(A.indices, dtype=np.int64)
        nzval = np.ascontiguousarray(A.data, dtype=np.float64)
        h = C.c_void_p()
        rc = lib.tlpb200_create(C.byref(h), self.m, self.n,
                                colptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                rowval.ctypes.data_as(C.POINTER(C.c_int64)),
                                _dp(nzval), 0, self._sys, C.byref(opt))
        self._h = h
        if rc != _lib.OK:
            try:
                _raise(rc, h)
            finally:
                lib.tlpb200_destroy(h)
                self._h = None

    # -- finalizer: the reference API has no close(); device memory is released here
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().tlpb200_destroy(h)
            except Exception:
                pass
            self._h = None

    close = __del__

    # -- KKT.update! --------------------------------------------------------------------------
    def update(self, theta_inv, regP, regD):
        theta_inv = np.ascontiguousarray(theta_inv, dtype=np.float64)
        regP = np.ascontiguousarray(regP, dtype=np.float64)
        regD = np.ascontiguousarray(regD, dtype=np.float64)
        # spd.jl:26-34 / sqd.jl:28-36
        if theta_inv.shape[0] != self.n:
            raise DimensionMismatch(f"length(θ)={theta_inv.shape[0]} but KKT solver has n={self.n}.")
        if regP.shape[0] != self.n:
            raise DimensionMismatch(f"length(regP)={regP.shape[0]} but KKT solver has n={self.n}")
        if regD.shape[0] != self.m:
            raise DimensionMismatch(f"length(regD)={regD.shape[0]} but KKT solver has m={self.m}")
        bad = C.c_int64(-1)
        rc = _lib.load().tlpb200_update(self._h, _dp(theta_inv), _dp(regP), _dp(regD), C.byref(bad))
        if rc != _lib.OK:
            _raise(rc, self._h)

    # -- KKT.solve! ---------------------------------------------------------------------------
    def solve(self, dx, dy, xi_p, xi_d):
        xi_p = np.ascontiguousarray(xi_p, dtype=np.float64)
        xi_d = np.ascontiguousarray(xi_d, dtype=np.float64)
        if dx.shape[0] != self.n or xi_d.shape[0] != self.n or dy.shape[0] != self.m or xi_p.shape[0] != self.m:
            raise DimensionMismatch("solve!: vector lengths do not match the KKT solver")
        if not (dx.flags.c_contiguous and dy.flags.c_contiguous and dx.dtype == np.float64 and dy.dtype == np.float64):
            raise TypeError("dx, dy must be contiguous float64 arrays (they are overwritten in place)")
        rc = _lib.load().tlpb200_solve(self._h, _dp(dx), _dp(dy), _dp(xi_p), _dp(xi_d), 1, self.n, self.m)
        if rc != _lib.OK:
            _raise(rc, self._h)

    def solve_multi(self, DX, DY, XI_P, XI_D):
        """nrhs right-hand sides, one per *row* of the (nrhs, n)/(nrhs, m) C-contiguous arrays."""
        nrhs = XI_P.shape[0]
        XI_P = np.ascontiguousarray(XI_P, dtype=np.float64); XI_D = np.ascontiguousarray(XI_D, dtype=np.float64)
        if XI_P.shape != (nrhs, self.m) or XI_D.shape != (nrhs, self.n) or DX.shape != (nrhs, self.n) or DY.shape != (nrhs, self.m):
            raise DimensionMismatch("solve_multi: expected (nrhs, m) / (nrhs, n) arrays")
        if not (DX.flags.c_contiguous and DY.flags.c_contiguous and DX.dtype == np.float64 and DY.dtype == np.float64):
            raise TypeError("DX, DY must be C-contiguous float64 arrays (they are overwritten in place)")
        rc = _lib.load().tlpb200_solve(self._h, _dp(DX), _dp(DY), _dp(np.ascontiguousarray(XI_P)),
                                       _dp(np.ascontiguousarray(XI_D)), nrhs, self.n, self.m)
        if rc != _lib.OK:
            _raise(rc, self._h)

    # -- device-resident variants (torch tensors on the solver's device) -------------------------
    def set_stream(self, cuda_stream_ptr):
        rc = _lib.load().tlpb200_set_stream(self._h, C.c_void_p(cuda_stream_ptr))
        if rc != _lib.OK:
            _raise(rc, self._h)

    def update_dev(self, theta_inv, regP, regD):
        rc = _lib.load().tlpb200_update_dev(self._h, C.c_void_p(theta_inv.data_ptr()), C.c_void_p(regP.data_ptr()),
                                            C.c_void_p(regD.data_ptr()))
        if rc != _lib.OK:
            _raise(rc, self._h)

    def update_status(self):
        bad = C.c_int64(-1)
        rc = _lib.load().tlpb200_update_status(self._h, C.byref(bad))
        if rc != _lib.OK:
            _raise(rc, self._h)

    def solve_dev(self, dx, dy, xi_p, xi_d):
        rc = _lib.load().tlpb200_solve_dev(self._h, C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()),
                                           C.c_void_p(xi_p.data_ptr()), C.c_void_p(xi_d.data_ptr()), 1, self.n, self.m)
        if rc != _lib.OK:
            _raise(rc, self._h)

    def solve_status(self):
        """synchronise and raise if a sweep kernel of the solve_dev calls enqueued so far timed out"""
        rc = _lib.load().tlpb200_solve_status(self._h)
        if rc != _lib.OK:
            _raise(rc, self._h)

    def debug_raise_timeout(self):
        rc = _lib.load().tlpb200_debug_raise_timeout(self._h)
        if rc != _lib.OK:
            _raise(rc, self._h)

    def synchronize(self):
        rc = _lib.load().tlpb200_synchronize(self._h)
        if rc != _lib.OK:
            _raise(rc, self._h)

    def set_profiling(self, on=True):
        rc = _lib.load().tlpb200_set_profiling(self._h, 1 if on else 0)
        if rc != _lib.OK:
            _raise(rc, self._h)

    # -- introspection ------------------------------------------------------------------------
    def stats(self):
        st = _lib.Stats()
        _lib.load().tlpb200_stats_get(self._h, C.byref(st))
        return st.asdict()

    def symbolic(self):
        st = self.stats()
        N, ns = st["order"], st["nsuper"]
        perm = np.zeros(N, np.int32); parent = np.zeros(N, np.int32); cc = np.zeros(N, np.int32)
        first = np.zeros(ns + 1, np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        _lib.load().tlpb200_get_symbolic(self._h, vp(perm), vp(parent), vp(cc), vp(first))
        rowptr = np.zeros(ns + 1, np.int64)
        _lib.load().tlpb200_get_structure(self._h, vp(rowptr), None)
        rows = np.zeros(int(rowptr[-1]), np.int32)
        _lib.load().tlpb200_get_structure(self._h, None, vp(rows))
        return dict(perm=perm, parent=parent, colcount=cc, sn_first=first, sn_rowptr=rowptr, sn_rows=rows)

    def dense_cols(self):
        """indices of the columns handled by the low-rank Schur correction (K1 only)"""
        cnt = C.c_int32(0)
        _lib.load().tlpb200_get_dense_cols(self._h, C.byref(cnt), None)
        ids = np.zeros(cnt.value, np.int64)
        if cnt.value:
            _lib.load().tlpb200_get_dense_cols(self._h, C.byref(cnt), ids.ctypes.data_as(C.POINTER(C.c_int64)))
        return ids

    def dist_info(self):
        """owner[s] (rank or -1 = replicated top part), offset and length of the top panels in Lx."""
        ns = self.stats()["nsuper"]
        owner = np.zeros(ns, np.int32)
        off = C.c_int64(0); cnt = C.c_int64(0)
        _lib.load().tlpb200_dist_info(self._h, C.c_void_p(owner.ctypes.data), C.byref(off), C.byref(cnt))
        return owner, off.value, cnt.value

    def debug_assembled(self, theta_inv, regP, regD):
        """Assemble only (no factorisation) and return (Lx, xptr) -- parity test of the assemble kernel."""
        lib = _lib.load()
        t = np.ascontiguousarray(theta_inv, np.float64); p = np.ascontiguousarray(regP, np.float64)
        d = np.ascontiguousarray(regD, np.float64)
        rc = lib.tlpb200_debug_assemble(self._h, _dp(t), _dp(p), _dp(d))
        if rc != _lib.OK:
            _raise(rc, self._h)
        return self.debug_lx()

    def debug_lx(self):
        lib = _lib.load()
        st = self.stats()
        lx = np.zeros(st["nnzL_stored"], np.float64)
        xptr = np.zeros(st["nsuper"] + 1, np.int64)
        rc = lib.tlpb200_debug_get_lx(self._h, _dp(lx), xptr.ctypes.data_as(C.POINTER(C.c_int64)))
        if rc != _lib.OK:
            _raise(rc, self._h)
        return lx, xptr


    PACK_DTYPE = np.dtype([("sn", "<i4"), ("r0", "<i4"), ("nr", "<i4"), ("j", "<i4"), ("fdst", "<i8"), ("bdst", "<i8")])
    TASK_DTYPE = np.dtype([("sn", "<i4"), ("kind", "<i4"), ("blk", "<i4"), ("r0", "<i4"), ("nr", "<i4"),
                           ("ntile", "<i4"), ("nbelow", "<i4"), ("xq0", "<i4"), ("tile0", "<i8")])

    def chain_times(self):
        """(forward, backward) globaltimer stamps [ns] of every block publish of the last dense sweeps
        (needs TLPB200_CHAIN_TIMES=1 in the environment at setup)."""
        lib = _lib.load()
        nb = C.c_int64(0)
        lib.tlpb200_debug_chain_times(self._h, None, C.byref(nb))
        out = np.zeros(2 * nb.value, np.uint64)
        rc = lib.tlpb200_debug_chain_times(self._h, C.c_void_p(out.ctypes.data), C.byref(nb))
        if rc != _lib.OK:
            _raise(rc, self._h)
        return out[:nb.value].astype(np.int64), out[nb.value:].astype(np.int64)

    def factor_trace(self):
        """[nlevels, 4, 2] globaltimer ns (first start, last end) of {diag, trsm, urgent, lazy} per level of the last
        update! (needs TLPB200_TRACE_FACTOR=1 in the environment at setup)."""
        lib = _lib.load()
        nl = C.c_int64(0)
        lib.tlpb200_debug_factor_trace(self._h, None, C.byref(nl))
        out = np.zeros((nl.value, 4, 2), np.uint64)
        rc = lib.tlpb200_debug_factor_trace(self._h, C.c_void_p(out.ctypes.data), C.byref(nl))
        if rc != _lib.OK:
            _raise(rc, self._h)
        return out.astype(np.int64)

    def update_plan(self):
        """Update-task plan (host data, available on analyze_only handles): FP64 tile tasks, tcgen05 tasks, pieces, views."""
        lib = _lib.load()
        cnt = np.zeros(10, np.int64)
        none9 = [None] * 9
        lib.tlpb200_debug_update_plan(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), *none9)
        upd = np.zeros((int(cnt[0]), 8), np.int32); upd128 = np.zeros((int(cnt[1]), 8), np.int32)
        oz = np.zeros((int(cnt[2]), 8), np.int32); pieces = np.zeros((int(cnt[3]), 4), np.int32)
        views = np.zeros((int(cnt[4]), 4), np.int32); panel = np.zeros((int(cnt[5]), 4), np.int32)
        levels = np.zeros((int(cnt[6]), int(cnt[7])), np.int32)
        small_list = np.zeros(int(cnt[8]), np.int32); level_pieces = np.zeros(int(cnt[9]), np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        lib.tlpb200_debug_update_plan(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), vp(upd), vp(upd128), vp(oz), vp(pieces), vp(views),
                                      vp(panel), vp(levels), vp(small_list), vp(level_pieces))
        return {"upd": upd, "upd128": upd128, "oz": oz, "pieces": pieces, "views": views, "panel": panel, "levels": levels,
                "small_list": small_list, "level_pieces": level_pieces}

    # field order of the int32 records of update_plan()["levels"] (LevelPlan in csrc/plan.hpp)
    LEVEL_FIELDS = ("small_begin", "small_end", "piece_begin", "piece_end", "panel_begin", "panel_end", "panel_crit_end",
                    "ext_begin", "ext_end", "ext_crit_end", "urgent_end", "lazy_begin", "lazy_end", "ext_atomic",
                    "fwd_begin", "fwd_end", "bwd_begin", "bwd_end", "fbig_begin", "fbig_end", "bbig_begin", "bbig_end",
                    "below_begin", "below_end", "oz_begin", "oz_end", "ozs_begin", "ozs_end", "oz_tile", "inv_end", "pack_end")

    ITEM_DTYPE = np.dtype([("sn", "<i4"), ("blk", "<i4"), ("kind", "<i4"), ("r0", "<i4"), ("nr", "<i4"), ("pad", "<i4")])

    def solve_ops(self):
        """Launch sequences of the triangular sweeps with merged levels (host data, available on analyze_only handles)."""
        lib = _lib.load()
        cnt = np.zeros(4, np.int64)
        none10 = [None] * 10
        lib.tlpb200_debug_solve_ops(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), *none10)
        ns = self.stats()["nsuper"]
        nbel = np.zeros(ns, np.int32)
        fo = np.zeros((int(cnt[0]), 4), np.int32); bo = np.zeros((int(cnt[1]), 4), np.int32)
        need = np.zeros(ns, np.int32); fpar = np.zeros(ns, np.int32); bwait = np.zeros(ns, np.int32); bn = np.zeros(ns, np.int32)
        fit = np.zeros(int(cnt[2]), self.ITEM_DTYPE); bit = np.zeros(int(cnt[3]), self.ITEM_DTYPE); par = np.zeros(ns, np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        lib.tlpb200_debug_solve_ops(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), vp(fo), vp(bo), vp(need), vp(fpar), vp(bwait),
                                    vp(bn), vp(fit), vp(bit), vp(par), vp(nbel))
        return dict(fwd_ops=fo, bwd_ops=bo, fwd_need=need, fwd_parent=fpar, bwd_wait=bwait, bwd_nitems=bn, fwd_items=fit,
                    bwd_seq=bit, sn_parent=par, bwd_nbelow=nbel)

    def phase_deps(self, phase):
        """sharded solver: (fwd_need, fwd_parent, bwd_wait) of the merged-level sweeps inside phase 0 (own subtrees) / 1 (top)"""
        ns = self.stats()["nsuper"]
        need = np.zeros(ns, np.int32); fpar = np.full(ns, -1, np.int32); bwait = np.full(ns, -1, np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        rc = _lib.load().tlpb200_debug_phase_deps(self._h, phase, vp(need), vp(fpar), vp(bwait))
        if rc != _lib.OK:
            _raise(rc, self._h)
        return need, fpar, bwait

    def big_plan(self):
        """Dense-solve plan of the big supernodes (host data, available on analyze_only handles)."""
        lib = _lib.load()
        cnt = np.zeros(6, np.int64)
        lib.tlpb200_debug_big_plan(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), None, None, None)
        pack = np.zeros(int(cnt[0]), self.PACK_DTYPE)
        fwd = np.zeros(int(cnt[1]), self.TASK_DTYPE)
        bwd = np.zeros(int(cnt[2]), self.TASK_DTYPE)
        lib.tlpb200_debug_big_plan(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64)), C.c_void_p(pack.ctypes.data),
                                   C.c_void_p(fwd.ctypes.data), C.c_void_p(bwd.ctypes.data))
        return {"pack": pack, "fwd": fwd, "bwd": bwd, "n_ftiles": int(cnt[3]), "n_btiles": int(cnt[4]),
                "xq_slots": int(cnt[5])}


# ---- the generic functions of src/KKT/KKT.jl ---------------------------------------------------
def setup(A, system=None, backend=None):
    """KKT.setup(A, system, backend) (KKT.jl:59)."""
    return B200KKTSolver(A, system if system is not None else DefaultKKTSystem(), backend or Backend())


def update_(kkt, theta_inv, regP, regD):
    """KKT.update!(kkt, θinv, regP, regD) (KKT.jl:83)."""
    kkt.update(theta_inv, regP, regD)


def solve_(dx, dy, kkt, xi_p, xi_d):
    """KKT.solve!(dx, dy, kkt, ξp, ξd) (KKT.jl:100)."""
    kkt.solve(dx, dy, xi_p, xi_d)


def arithmetic(kkt):
    """KKT.arithmetic(kkt) (KKT.jl:107)."""
    return np.float64


def backend(kkt):
    """KKT.backend(kkt) (KKT.jl:114)."""
    return _lib.load().tlpb200_backend_name().decode()


def linear_system(kkt):
    """KKT.linear_system(kkt) (KKT.jl:121)."""
    return _lib.load().tlpb200_linear_system(kkt._h).decode()
