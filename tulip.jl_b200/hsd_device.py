"""Device-resident HSD driver: the Python face of ``tlpb200_hsd_*`` (include/tlpb200.h).

Same algorithm and options as the host mirror ``hsd.HSD`` (<- /root/reference/src/IPM/HSD/HSD.jl:203-350,
src/IPM/HSD/step.jl:10-401), but the iterate, the residuals and every work vector of ``compute_step!`` live in HBM and the
theta / rhs / recovery / step-length / corrector-target passes are fused CUDA kernels (csrc/kernels_ipm.cu); the host reads
back one block of scalars per control-flow decision.  SURVEY 8f-1 / 8f-2.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .hsd import IPMOptions
from .kkt import B200KKTSolver, _dp, _raise


class DeviceHSD:
    """``HSD(dat, kkt_options)`` (HSD.jl:34-63) on the device; ``kkt`` is a B200KKTSolver (or DistB200KKT.local)."""

    def __init__(self, kkt: B200KKTSolver, b, c, l, u, c0=0.0, params: IPMOptions | None = None):
        self.kkt = kkt
        self.m, self.n = kkt.m, kkt.n
        self.params = params or IPMOptions()
        b = np.ascontiguousarray(b, np.float64); c = np.ascontiguousarray(c, np.float64)
        l = np.ascontiguousarray(l, np.float64); u = np.ascontiguousarray(u, np.float64)
        if b.shape[0] != self.m or c.shape[0] != self.n or l.shape[0] != self.n or u.shape[0] != self.n:
            raise ValueError("b, c, l, u do not match the KKT solver's A")
        rc = _lib.load().tlpb200_hsd_create(kkt._h, _dp(b), _dp(c), _dp(l), _dp(u), float(c0))
        if rc != _lib.OK:
            _raise(rc, kkt._h)
        self.info = None
        self.status = "Trm_Unknown"
        self.niter = 0

    def _options(self):
        P = self.params
        o = _lib.HSDOptions()
        _lib.load().tlpb200_hsd_default_options(C.byref(o))
        o.iterations_limit = int(P.IterationsLimit); o.correction_limit = int(P.CorrectionLimit)
        o.time_limit = float(P.TimeLimit)
        o.tol_pfeas, o.tol_dfeas, o.tol_rgap, o.tol_ifeas = P.TolerancePFeas, P.ToleranceDFeas, P.ToleranceRGap, P.ToleranceIFeas
        o.step_damp, o.gamma_min, o.centrality_outlier = P.StepDampFactor, P.GammaMin, P.CentralityOutlierThreshold
        o.preg_min, o.dreg_min = P.PRegMin, P.DRegMin
        return o

    def _take(self, info):
        self.info = {k: getattr(info, k) for k, _ in info._fields_}
        self.status = _lib.TRM_STATUS[info.status]
        self.niter = int(info.niter)
        self.primal_objective = float(info.pobj); self.dual_objective = float(info.dobj)
        self.n_update = int(info.n_update); self.n_solve = int(info.n_solve)
        self.t_factor = float(info.ms_update) * 1e-3; self.t_solve = float(info.ms_solve) * 1e-3

    def reset(self):
        rc = _lib.load().tlpb200_hsd_reset(self.kkt._h)
        if rc != _lib.OK:
            _raise(rc, self.kkt._h)

    def iterate(self):
        """one pass of the main loop body; returns the status string ("Trm_Unknown" = go on)"""
        o = self._options(); info = _lib.HSDInfo()
        rc = _lib.load().tlpb200_hsd_iterate(self.kkt._h, C.byref(o), C.byref(info))
        self._take(info)
        if rc != _lib.OK:
            _raise(rc, self.kkt._h)
        return self.status

    def optimize(self):
        o = self._options(); info = _lib.HSDInfo()
        rc = _lib.load().tlpb200_hsd_optimize(self.kkt._h, C.byref(o), C.byref(info))
        self._take(info)
        if rc != _lib.OK:
            _raise(rc, self.kkt._h)
        return self.status

    @property
    def log(self):
        """rows (iter, pobj, dobj, pfeas, dfeas, gfeas, mu) like hsd.HSD.log"""
        rows = C.c_int64(0)
        lib = _lib.load()
        lib.tlpb200_hsd_get_log(self.kkt._h, None, C.byref(rows))
        out = np.zeros((rows.value, 8))
        if rows.value:
            lib.tlpb200_hsd_get_log(self.kkt._h, _dp(out), C.byref(rows))
        return [tuple([int(r[0])] + [float(v) for v in r[1:7]]) for r in out]

    def point(self):
        n, m = self.n, self.m
        x = np.zeros(n); xl = np.zeros(n); xu = np.zeros(n); y = np.zeros(m); zl = np.zeros(n); zu = np.zeros(n); tk = np.zeros(2)
        rc = _lib.load().tlpb200_hsd_get_point(self.kkt._h, _dp(x), _dp(xl), _dp(xu), _dp(y), _dp(zl), _dp(zu), _dp(tk))
        if rc != _lib.OK:
            _raise(rc, self.kkt._h)
        return dict(x=x, xl=xl, xu=xu, y=y, zl=zl, zu=zu, tau=float(tk[0]), kappa=float(tk[1]))
