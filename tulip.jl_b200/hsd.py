"""Host-side mirror of the *caller* of the KKT hot path: Tulip's homogeneous self-dual IPM.

Python twin of /root/reference/src/IPM/HSD/HSD.jl (main loop :203-350, residuals :77-128, status
:136-196) and src/IPM/HSD/step.jl (compute_step! :10-151, solve_newton_system! :198-266,
max_step_length :274-306, compute_higher_corrector! :325-401), written so that any object with
``update(θinv, regP, regD)`` / ``solve(dx, dy, ξp, ξd)`` -- in particular ``B200KKTSolver`` -- is
driven exactly the way ``HSD`` drives ``hsd.kkt``.  It is what the benchmark runs to produce real
(θ, regP, regD, ξ) sequences and to count IPM iterations; it is NOT the oracle (that is
oracle/hsd_ref.py, used by the tests to check this file's trajectory).

Timing follows the reference's TimerOutputs sections: "Factorization" = Σ update! (step.jl:37),
"KKT" = Σ solve! (step.jl:63, :214).
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

_SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


@dataclass
class IPMOptions:                      # src/IPM/options.jl:1-25
    OutputLevel: int = 0
    IterationsLimit: int = 100
    TimeLimit: float = float("inf")
    TolerancePFeas: float = _SQRT_EPS
    ToleranceDFeas: float = _SQRT_EPS
    ToleranceRGap: float = _SQRT_EPS
    ToleranceIFeas: float = _SQRT_EPS
    CorrectionLimit: int = 3
    StepDampFactor: float = 0.9995
    GammaMin: float = 0.1
    CentralityOutlierThreshold: float = 0.1
    PRegMin: float = _SQRT_EPS
    DRegMin: float = _SQRT_EPS
    Threads: int = 1                   # src/parameters.jl:7 -- model.jl:73 BLAS.set_num_threads(params.Threads)


class _Dir:
    __slots__ = ("x", "xl", "xu", "y", "zl", "zu", "tau", "kappa")

    def __init__(self, m, n):
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.y = np.zeros(m); self.zl = np.zeros(n); self.zu = np.zeros(n)
        self.tau = 0.0; self.kappa = 0.0


def _max_step(x, dx):                   # step.jl:274-288
    neg = dx < 0.0
    return float(np.min(-x[neg] / dx[neg])) if neg.any() else np.inf


class HSD:
    """HSD(dat, kkt_options) with the KKT solver injected (HSD.jl:34-63)."""

    def __init__(self, A, b, c, l, u, kkt, c0=0.0, objsense=True, params: IPMOptions | None = None):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.AT = self.A.T.tocsc()
        self.m, self.n = self.A.shape
        self.b = np.asarray(b, float); self.c = np.asarray(c, float)
        self.l = np.asarray(l, float); self.u = np.asarray(u, float)
        self.c0 = float(c0); self.objsense = bool(objsense)
        self.lf = np.isfinite(self.l); self.uf = np.isfinite(self.u)      # ipmdata.jl:44-45
        self.lm = np.where(self.lf, self.l, 0.0); self.um = np.where(self.uf, self.u, 0.0)
        self.p = int(self.lf.sum() + self.uf.sum())                      # HSD.jl:39
        self.kkt = kkt
        self.params = params or IPMOptions()
        m, n = self.m, self.n
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.y = np.zeros(m); self.zl = np.zeros(n); self.zu = np.zeros(n)
        self.tau = 1.0; self.kappa = 1.0; self.mu = 1.0
        self.regP = np.ones(n); self.regD = np.ones(m); self.regG = 1.0   # HSD.jl:50-52
        self.niter = 0
        self.status = "Trm_Unknown"
        self.primal_objective = np.inf; self.dual_objective = -np.inf
        self.t_factor = 0.0; self.t_solve = 0.0; self.n_update = 0; self.n_solve = 0
        self.solves_per_iter = []
        self.log = []
        self.on_update = None    # optional hook(theta_inv, regP, regD) -- lets the benchmark record inputs

    # point.jl:45-48
    def _update_mu(self):
        self.mu = (self.xl @ self.zl + self.xu @ self.zu + self.tau * self.kappa) / (self.p + 1)

    # HSD.jl:77-128
    def compute_residuals(self):
        self.rp = self.tau * self.b - self.A @ self.x
        self.rl = np.where(self.lf, -self.x + self.xl + self.tau * self.lm, 0.0)
        self.ru = np.where(self.uf, -self.x - self.xu + self.tau * self.um, 0.0)
        self.rd = self.tau * self.c - self.AT @ self.y + np.where(self.uf, self.zu, 0.0) - np.where(self.lf, self.zl, 0.0)
        dual = self.b @ self.y + self.lm @ self.zl - self.um @ self.zu
        cx = self.c @ self.x
        self.rg = self.kappa + (cx - dual)
        ninf = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
        self.rp_nrm, self.rl_nrm, self.ru_nrm, self.rd_nrm = ninf(self.rp), ninf(self.rl), ninf(self.ru), ninf(self.rd)
        self.rg_nrm = abs(self.rg)
        self.primal_objective = cx / self.tau + self.c0
        self.dual_objective = dual / self.tau + self.c0

    # HSD.jl:136-196
    def update_solver_status(self):
        P = self.params
        ninf = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
        nb, nl, nu, nc = ninf(self.b), ninf(self.lm), ninf(self.um), ninf(self.c)
        self.status = "Trm_Unknown"
        rho_p = max(self.rp_nrm / (self.tau * (1 + nb)), self.rl_nrm / (self.tau * (1 + nl)),
                    self.ru_nrm / (self.tau * (1 + nu)))
        rho_d = self.rd_nrm / (self.tau * (1 + nc))
        rho_g = abs(self.primal_objective - self.dual_objective) / (1 + abs(self.dual_objective))
        if rho_p <= P.TolerancePFeas and rho_d <= P.ToleranceDFeas and rho_g <= P.ToleranceRGap:
            self.status = "Trm_Optimal"
            return
        cx = self.c @ self.x
        lhs = max(ninf(self.A @ self.x), ninf(np.where(self.lf, self.x - self.xl, 0.0)),
                  ninf(np.where(self.uf, self.x + self.xu, 0.0))) * (nc / max(1.0, nb))
        if lhs < -P.ToleranceIFeas * cx:
            self.status = "Trm_DualInfeasible"
            return
        delta = self.AT @ self.y + np.where(self.lf, self.zl, 0.0) - np.where(self.uf, self.zu, 0.0)
        dualobj = self.b @ self.y + self.lm @ self.zl - self.um @ self.zu
        if ninf(delta) * max(nl, nu, nb) / max(1.0, nc) < dualobj * P.ToleranceIFeas:
            self.status = "Trm_PrimalInfeasible"

    # HSD.jl:203-350
    def optimize(self, max_iter=None, callback=None):
        # model.jl:73: BLAS.set_num_threads(params.Threads) (default 1).  Besides mirroring the reference this
        # matters on the GPU box: a multi-threaded ddot leaves 100+ OpenBLAS workers spinning, which starves the
        # CUDA driver threads and more than doubles the wall time of the next update! (measured).
        try:
            from threadpoolctl import threadpool_limits
            ctx = threadpool_limits(limits=max(1, int(self.params.Threads)), user_api="blas")
        except Exception:          # pragma: no cover
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            return self._optimize(max_iter, callback)

    def _optimize(self, max_iter=None, callback=None):
        P = self.params
        tstart = time.time()
        self.niter = 0
        self.x[:] = 0.0; self.y[:] = 0.0                                 # HSD.jl:238-247
        self.xl[:] = self.lf; self.xu[:] = self.uf
        self.zl[:] = self.lf; self.zu[:] = self.uf
        self.tau = 1.0; self.kappa = 1.0
        self._update_mu()
        limit = P.IterationsLimit if max_iter is None else max_iter
        while True:
            self.compute_residuals()
            self._update_mu()
            self.log.append((self.niter, self.primal_objective, self.dual_objective,
                             max(self.rp_nrm, self.ru_nrm), self.rd_nrm, self.rg_nrm, self.mu))
            if P.OutputLevel > 0:
                sgn = 1.0 if self.objsense else -1.0
                print("%4d  %+14.7e  %+14.7e  %8.2e %8.2e %8.2e  %7.1e  %.2f" % (
                    self.niter, sgn * self.primal_objective, sgn * self.dual_objective,
                    max(self.rp_nrm, self.ru_nrm), self.rd_nrm, self.rg_nrm, self.mu, time.time() - tstart), flush=True)
            self.update_solver_status()
            if self.status in ("Trm_Optimal", "Trm_PrimalInfeasible", "Trm_DualInfeasible"):
                break
            if self.niter >= limit:
                self.status = "Trm_IterationLimit"
                break
            if time.time() - tstart >= P.TimeLimit:
                self.status = "Trm_TimeLimit"
                break
            try:
                self.compute_step()
            except Exception as err:                                     # HSD.jl:321-339
                nm = type(err).__name__
                if nm in ("PosDefException", "SingularException"):
                    self.status = "Trm_NumericalProblem"
                elif nm in ("OutOfMemoryError", "MemoryError"):
                    self.status = "Trm_MemoryLimit"
                else:
                    raise
                break
            self.niter += 1
            if callback is not None:
                callback(self)
        return self.status

    def _solve(self, dx, dy, xi_p, xi_d):
        t0 = time.perf_counter()
        self.kkt.solve(dx, dy, xi_p, xi_d)
        self.t_solve += time.perf_counter() - t0
        self.n_solve += 1

    # step.jl:10-151
    def compute_step(self):
        P = self.params
        m, n = self.m, self.n
        with np.errstate(divide="ignore", invalid="ignore"):
            self.ixl = np.where(self.lf, 1.0 / self.xl, 0.0)
            self.ixu = np.where(self.uf, 1.0 / self.xu, 0.0)
        thl = self.zl * self.ixl                                         # step.jl:24-26
        thu = self.zu * self.ixu
        thinv = thl + thu
        self.regP = np.maximum(P.PRegMin, self.regP / 10)                # step.jl:29-31
        self.regD = np.maximum(P.DRegMin, self.regD / 10)
        self.regG = max(P.PRegMin, self.regG / 10)
        n_solve0 = self.n_solve
        nbump = 0
        while nbump <= 3:                                                # step.jl:34-51
            if self.on_update is not None:
                self.on_update(thinv, self.regP, self.regD)
            t0 = time.perf_counter()
            try:
                self.kkt.update(thinv, self.regP, self.regD)
                self.t_factor += time.perf_counter() - t0
                self.n_update += 1
                break
            except Exception as err:
                self.t_factor += time.perf_counter() - t0
                self.n_update += 1
                if type(err).__name__ not in ("PosDefException", "ZeroPivotException"):
                    raise
                self.regD = self.regD * 100; self.regP = self.regP * 100; self.regG *= 100
                nbump += 1
        if not nbump < 3:
            from .kkt import PosDefException
            raise PosDefException("factorization could not be saved")    # step.jl:51

        hx = np.zeros(n); hy = np.zeros(m)
        cbar = self.c + thl * self.lm + thu * self.um
        self._solve(hx, hy, self.b, self.c - thl * self.lm - thu * self.um)   # step.jl:61-63
        h0 = (self.lm @ (self.lm * thl) + self.um @ (self.um * thu) - cbar @ hx + self.b @ hy
              + self.kappa / self.tau + self.regG)                       # step.jl:69-76
        self._cbar, self._thl, self._thu = cbar, thl, thu

        D = _Dir(m, n)
        self._newton(D, hx, hy, h0, self.rp, self.rl, self.ru, self.rd, self.rg,
                     np.where(self.lf, -self.xl * self.zl, 0.0), np.where(self.uf, -self.xu * self.zu, 0.0),
                     -self.tau * self.kappa)                             # step.jl:79-85
        alpha = self._alpha(D)
        gamma = (1 - alpha) ** 2 * min(1 - alpha, P.GammaMin)            # step.jl:89-90
        eta = 1.0 - gamma
        self._newton(D, hx, hy, h0, eta * self.rp, eta * self.rl, eta * self.ru, eta * self.rd, eta * self.rg,
                     np.where(self.lf, -self.xl * self.zl + gamma * self.mu - D.xl * D.zl, 0.0),
                     np.where(self.uf, -self.xu * self.zu + gamma * self.mu - D.xu * D.zu, 0.0),
                     -self.tau * self.kappa + gamma * self.mu - D.tau * D.kappa)   # step.jl:93-99
        alpha = self._alpha(D)
        ncor = 0
        while ncor < P.CorrectionLimit and alpha < 0.999:                # step.jl:104-136
            a_prev = alpha
            ncor += 1
            Dc = _Dir(m, n)
            a_c = self._corrector(Dc, gamma, hx, hy, h0, D, a_prev, P.CentralityOutlierThreshold)
            if a_c > a_prev:
                D = Dc
                alpha = a_c
            if a_c < 1.1 * a_prev:
                break
        alpha *= P.StepDampFactor                                        # step.jl:139-148
        self.x += alpha * D.x; self.xl += alpha * D.xl; self.xu += alpha * D.xu
        self.y += alpha * D.y; self.zl += alpha * D.zl; self.zu += alpha * D.zu
        self.tau += alpha * D.tau; self.kappa += alpha * D.kappa
        self._update_mu()
        self.solves_per_iter.append(self.n_solve - n_solve0)

    def _alpha(self, D):                                                 # step.jl:295-306
        at = (-self.tau / D.tau) if D.tau < 0 else 1.0
        ak = (-self.kappa / D.kappa) if D.kappa < 0 else 1.0
        return min(1.0, _max_step(self.xl, D.xl), _max_step(self.xu, D.xu), _max_step(self.zl, D.zl),
                   _max_step(self.zu, D.zu), at, ak)

    # step.jl:198-266
    def _newton(self, D, hx, hy, h0, xi_p, xi_l, xi_u, xi_d, xi_g, xi_xzl, xi_xzu, xi_tk):
        ixl, ixu = self.ixl, self.ixu
        xi_d_ = xi_d - (xi_xzl + self.zl * xi_l) * ixl + (xi_xzu - self.zu * xi_u) * ixu   # step.jl:210-213
        self._solve(D.x, D.y, xi_p, xi_d_)                               # step.jl:214
        xi_g_ = (xi_g + xi_tk / self.tau - (xi_xzl * ixl) @ self.lm + (xi_xzu * ixu) @ self.um
                 - (self._thl * xi_l) @ self.lm - (self._thu * xi_u) @ self.um)          # step.jl:218-223
        D.tau = (xi_g_ + self._cbar @ D.x - self.b @ D.y) / h0           # step.jl:225-232
        D.x += D.tau * hx
        D.y += D.tau * hy
        D.xl = np.where(self.lf, -xi_l + D.x - D.tau * self.lm, 0.0)     # step.jl:240-245
        D.xu = np.where(self.uf, xi_u - D.x + D.tau * self.um, 0.0)
        D.zl = (xi_xzl - self.zl * D.xl) * ixl                           # step.jl:248-249
        D.zu = (xi_xzu - self.zu * D.xu) * ixu
        D.kappa = (xi_tk - self.kappa * D.tau) / self.tau                # step.jl:252

    # step.jl:325-401
    def _corrector(self, Dc, gamma, hx, hy, h0, D, alpha, beta):
        a_ = min(1.0, 2.0 * alpha)
        vl = np.where(self.lf, (self.xl + a_ * D.xl) * (self.zl + a_ * D.zl), 0.0)
        vu = np.where(self.uf, (self.xu + a_ * D.xu) * (self.zu + a_ * D.zu), 0.0)
        vt = (self.tau + a_ * D.tau) * (self.kappa + a_ * D.kappa)
        mu_l = beta * self.mu * gamma
        mu_u = gamma * self.mu / beta

        def tgt(v, flag):
            return np.where(flag, np.where(v < mu_l, mu_l - v, np.where(v > mu_u, mu_u - v, 0.0)), v)
        vl = tgt(vl, self.lf); vu = tgt(vu, self.uf)
        vt = (mu_l - vt) if vt < mu_l else ((mu_u - vt) if vt > mu_u else 0.0)
        delta = (vl.sum() + vu.sum() + vt) / (self.p + 1)                # step.jl:374-377
        vl = vl - delta; vu = vu - delta; vt -= delta
        z_m = np.zeros(self.m); z_n = np.zeros(self.n)
        self._newton(Dc, hx, hy, h0, z_m, z_n, z_n, z_n, 0.0, vl, vu, vt)
        Dc.x += D.x; Dc.xl += D.xl; Dc.xu += D.xu; Dc.y += D.y
        Dc.zl += D.zl; Dc.zu += D.zu; Dc.tau += D.tau; Dc.kappa += D.kappa
        return self._alpha(Dc)
