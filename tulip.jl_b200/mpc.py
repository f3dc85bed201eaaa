"""Host-side mirror of the reference's SECOND client of the KKT boundary: Mehrotra's predictor-corrector IPM
(/root/reference/src/IPM/MPC/MPC.jl: residuals :101-141, status tests :149-211, main loop :218-351, starting point
:353-410; src/IPM/MPC/step.jl: compute_step! :10-123, solve_newton_system! :164-206, max_step_length_pd :213-223,
predictor / corrector / extra corrections :229-322, compute_target! :329-358).

Any object with ``update(theta_inv, regP, regD)`` / ``solve(dx, dy, xi_p, xi_d)`` is driven exactly the way ``MPC`` drives
``mpc.kkt``: one ``update!`` + two ``solve!`` for the starting point (MPC.jl:359-363), then per iteration one ``update!``
and 2 + <= CorrectionLimit ``solve!`` calls (no (hx, hy) solve: the formulation is not homogeneous).  SURVEY 8f-4: the
B200 backend must work under this caller too.  Like hsd.py this is the driver of tests / benchmarks, not a KKT kernel.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp

from .hsd import IPMOptions, _max_step

_SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


class _Dir:
    __slots__ = ("x", "xl", "xu", "y", "zl", "zu")

    def __init__(self, m, n):
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.y = np.zeros(m); self.zl = np.zeros(n); self.zu = np.zeros(n)

    def copy_from(self, o):
        for k in self.__slots__:
            getattr(self, k)[:] = getattr(o, k)


class MPC:
    """MPC(dat, kkt_options) with the KKT solver injected (MPC.jl:50-92)."""

    def __init__(self, A, b, c, l, u, kkt, c0=0.0, objsense=True, params: IPMOptions | None = None):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.AT = self.A.T.tocsc()
        self.m, self.n = self.A.shape
        self.b = np.asarray(b, float); self.c = np.asarray(c, float)
        self.l = np.asarray(l, float); self.u = np.asarray(u, float)
        self.c0 = float(c0); self.objsense = bool(objsense)
        self.lf = np.isfinite(self.l); self.uf = np.isfinite(self.u)
        self.lm = np.where(self.lf, self.l, 0.0); self.um = np.where(self.uf, self.u, 0.0)
        self.p = int(self.lf.sum() + self.uf.sum())                      # MPC.jl:55
        self.kkt = kkt
        self.params = params or IPMOptions()
        m, n = self.m, self.n
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.y = np.zeros(m); self.zl = np.zeros(n); self.zu = np.zeros(n)
        self.tau = 1.0; self.kappa = 0.0; self.mu = 1.0
        self.regP = np.ones(n); self.regD = np.ones(m)                   # MPC.jl:76-77
        self.alpha_p = 0.0; self.alpha_d = 0.0
        self.niter = 0
        self.status = "Trm_Unknown"
        self.primal_objective = np.inf; self.dual_objective = -np.inf
        self.t_factor = 0.0; self.t_solve = 0.0; self.n_update = 0; self.n_solve = 0
        self.log = []

    # -- timed KKT calls ("Factorization" / "KKT" sections of the reference's timer) --------------
    def _update(self, th, rp, rd):
        t0 = time.perf_counter()
        try:
            self.kkt.update(th, rp, rd)
        finally:
            self.t_factor += time.perf_counter() - t0
            self.n_update += 1

    def _solve(self, dx, dy, xi_p, xi_d):
        t0 = time.perf_counter()
        self.kkt.solve(dx, dy, xi_p, xi_d)
        self.t_solve += time.perf_counter() - t0
        self.n_solve += 1

    def _update_mu(self):                                               # point.jl:45-48 with hflag = false
        self.mu = (self.xl @ self.zl + self.xu @ self.zu) / self.p if self.p else 0.0

    # MPC.jl:101-141
    def compute_residuals(self):
        self.rp = self.b - self.A @ self.x
        self.rl = np.where(self.lf, (self.lm + self.xl) - self.x, 0.0)
        self.ru = np.where(self.uf, self.um - (self.x + self.xu), 0.0)
        self.rd = self.c - self.AT @ self.y + np.where(self.uf, self.zu, 0.0) - np.where(self.lf, self.zl, 0.0)
        ninf = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
        self.rp_nrm, self.rl_nrm, self.ru_nrm, self.rd_nrm = ninf(self.rp), ninf(self.rl), ninf(self.ru), ninf(self.rd)
        self.primal_objective = self.c @ self.x + self.c0
        self.dual_objective = self.b @ self.y + self.lm @ self.zl - self.um @ self.zu + self.c0

    # MPC.jl:149-211
    def update_solver_status(self):
        P = self.params
        ninf = lambda v: float(np.max(np.abs(v))) if v.size else 0.0
        nb, nl, nu, nc = ninf(self.b), ninf(self.lm), ninf(self.um), ninf(self.c)
        self.status = "Trm_Unknown"
        rho_p = max(self.rp_nrm / (1 + nb), self.rl_nrm / (1 + nl), self.ru_nrm / (1 + nu))
        rho_d = self.rd_nrm / (1 + nc)
        rho_g = abs(self.primal_objective - self.dual_objective) / (1 + abs(self.primal_objective))
        if rho_p <= P.TolerancePFeas and rho_d <= P.ToleranceDFeas and rho_g <= P.ToleranceRGap:
            self.status = "Trm_Optimal"
            return
        lhs = max(ninf(self.A @ self.x), ninf(np.where(self.lf, self.x - self.xl, 0.0)),
                  ninf(np.where(self.uf, self.x + self.xu, 0.0))) * (nc / max(1.0, nb))
        if lhs < -P.ToleranceIFeas * (self.c @ self.x):
            self.status = "Trm_DualInfeasible"
            return
        delta = self.AT @ self.y + np.where(self.lf, self.zl, 0.0) - np.where(self.uf, self.zu, 0.0)
        dualobj = self.b @ self.y + self.lm @ self.zl - self.um @ self.zu
        if ninf(delta) * max(nl, nu, nb) / max(1.0, nc) < dualobj * P.ToleranceIFeas:
            self.status = "Trm_PrimalInfeasible"

    # MPC.jl:353-410
    def compute_starting_point(self):
        m, n = self.m, self.n
        self._update(np.zeros(n), np.ones(n), 1e-6 * np.ones(m))        # MPC.jl:359
        self._solve(np.zeros(n), self.y, np.zeros(m), self.c)           # :362  for y
        self._solve(self.x, np.zeros(m), self.b, np.zeros(n))           # :363  for x
        lo = np.where(self.lf, self.x - self.lm, 0.0); hi = np.where(self.uf, self.um - self.x, 0.0)
        dx = 1.0 + max(0.0, -1.5 * (lo.min() if n else 0.0), -1.5 * (hi.min() if n else 0.0))
        self.xl = np.where(self.lf, (self.x - self.lm) + dx, 0.0)
        self.xu = np.where(self.uf, (self.um - self.x) + dx, 0.0)
        z = self.c - self.AT @ self.y
        cnt = self.lf.astype(float) + self.uf.astype(float)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.zl = np.where(self.lf, z / cnt, 0.0)
            self.zu = np.where(self.uf, -z / cnt, 0.0)
        dz = 1.0 + max(0.0, -1.5 * (self.zl.min() if n else 0.0), -1.5 * (self.zu.min() if n else 0.0))
        self.zl[self.lf] += dz
        self.zu[self.uf] += dz
        self.tau = 1.0; self.kappa = 0.0
        mu = self.xl @ self.zl + self.xu @ self.zu                      # :398-405 balance the complementarity products
        ddx = mu / (2 * (self.zl.sum() + self.zu.sum()))
        ddz = mu / (2 * (self.xl.sum() + self.xu.sum()))
        self.xl[self.lf] += ddx; self.xu[self.uf] += ddx
        self.zl[self.lf] += ddz; self.zu[self.uf] += ddz
        self._update_mu()

    # MPC.jl:218-351
    def optimize(self, max_iter=None, callback=None):
        try:
            from threadpoolctl import threadpool_limits
            ctx = threadpool_limits(limits=max(1, int(self.params.Threads)), user_api="blas")   # model.jl:73
        except Exception:          # pragma: no cover
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            return self._optimize(max_iter, callback)

    def _optimize(self, max_iter, callback):
        P = self.params
        tstart = time.time()
        self.niter = 0
        self.compute_starting_point()
        limit = P.IterationsLimit if max_iter is None else max_iter
        while True:
            self.compute_residuals()
            self._update_mu()
            self.log.append((self.niter, self.primal_objective, self.dual_objective,
                             max(self.rp_nrm, self.rl_nrm, self.ru_nrm), self.rd_nrm, 0.0, self.mu))
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
            except Exception as err:                                     # MPC.jl:320-340
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

    # step.jl:10-123
    def compute_step(self):
        P = self.params
        m, n = self.m, self.n
        with np.errstate(divide="ignore", invalid="ignore"):
            thl = np.where(self.lf, self.zl / self.xl, 0.0)             # step.jl:24-26
            thu = np.where(self.uf, self.zu / self.xu, 0.0)
        thinv = thl + thu
        self.regP = np.clip(self.regP / 10, _SQRT_EPS, 1.0)             # step.jl:29-32
        self.regD = np.clip(self.regD / 10, _SQRT_EPS, 1.0)
        nbump = 0
        while nbump <= 3:                                                # step.jl:35-51
            try:
                self._update(thinv, self.regP, self.regD)
                break
            except Exception as err:
                if type(err).__name__ not in ("PosDefException", "ZeroPivotException"):
                    raise
                self.regD = self.regD * 100; self.regP = self.regP * 100
                nbump += 1
        if not nbump < 3:
            from .kkt import PosDefException
            raise PosDefException("factorization could not be saved")    # step.jl:53
        D = _Dir(m, n); Dc = _Dir(m, n)
        # predictor (step.jl:229-246)
        xi_p, xi_l, xi_u, xi_d = self.rp.copy(), self.rl.copy(), self.ru.copy(), self.rd.copy()
        xzl = np.where(self.lf, -(self.xl * self.zl), 0.0)
        xzu = np.where(self.uf, -(self.xu * self.zu), 0.0)
        self._newton(D, xi_p, xi_l, xi_u, xi_d, xzl, xzu)
        self.alpha_p, self.alpha_d = self._alpha_pd(D)
        # corrector (step.jl:251-277)
        ap, ad = self.alpha_p, self.alpha_d
        mu_a = (np.where(self.lf, self.xl + ap * D.xl, 0.0) @ (self.zl + ad * D.zl)
                + np.where(self.uf, self.xu + ap * D.xu, 0.0) @ (self.zu + ad * D.zu)) / self.p
        sigma = float(np.clip((mu_a / self.mu) ** 3, _SQRT_EPS, 1.0 - _SQRT_EPS))
        xzl = np.where(self.lf, sigma * self.mu - D.xl * D.zl - self.xl * self.zl, 0.0)
        xzu = np.where(self.uf, sigma * self.mu - D.xu * D.zu - self.xu * self.zu, 0.0)
        self._newton(Dc, xi_p, xi_l, xi_u, xi_d, xzl, xzu)
        self.alpha_p, self.alpha_d = self._alpha_pd(Dc)
        D.copy_from(Dc)
        # extra centrality corrections (step.jl:73-107, :282-322)
        ncor = 0
        z_m = np.zeros(m); z_n = np.zeros(n)
        while ncor < P.CorrectionLimit:
            self._extra_correction(D, Dc, z_m, z_n)
            apc, adc = self._alpha_pd(Dc)
            if apc >= 1.01 * self.alpha_p and adc >= 1.01 * self.alpha_d:
                self.alpha_p, self.alpha_d = apc, adc
                D.copy_from(Dc)
                ncor += 1
            else:
                break
        self.alpha_p *= P.StepDampFactor                                 # step.jl:110-119
        self.alpha_d *= P.StepDampFactor
        self.x += self.alpha_p * D.x; self.xl += self.alpha_p * D.xl; self.xu += self.alpha_p * D.xu
        self.y += self.alpha_d * D.y; self.zl += self.alpha_d * D.zl; self.zu += self.alpha_d * D.zu
        self._update_mu()

    # step.jl:164-206
    def _newton(self, D, xi_p, xi_l, xi_u, xi_d, xzl, xzu):
        with np.errstate(divide="ignore", invalid="ignore"):
            xi_d_ = xi_d + np.where(self.lf, -((xzl + self.zl * xi_l) / self.xl), 0.0) \
                + np.where(self.uf, (xzu - self.zu * xi_u) / self.xu, 0.0)
            self._solve(D.x, D.y, xi_p, xi_d_)
            D.xl = np.where(self.lf, -xi_l + D.x, 0.0)
            D.xu = np.where(self.uf, xi_u - D.x, 0.0)
            D.zl = np.where(self.lf, (xzl - self.zl * D.xl) / self.xl, 0.0)
            D.zu = np.where(self.uf, (xzu - self.zu * D.xu) / self.xu, 0.0)

    def _alpha_pd(self, D):                                              # step.jl:213-223
        ap = min(1.0, _max_step(self.xl, D.xl), _max_step(self.xu, D.xu))
        ad = min(1.0, _max_step(self.zl, D.zl), _max_step(self.zu, D.zu))
        return ap, ad

    # step.jl:282-322 + compute_target! :329-358
    def _extra_correction(self, D, Dc, z_m, z_n, delta=0.3, gamma=0.1):
        ap, ad = self.alpha_p, self.alpha_d
        ap_ = min(ap + delta, 1.0); ad_ = min(ad + delta, 1.0)
        g = self.xl @ self.zl + self.xu @ self.zu
        ga = (np.where(self.lf, self.xl + ap * D.xl, 0.0) @ (self.zl + ad * D.zl)
              + np.where(self.uf, self.xu + ap * D.xu, 0.0) @ (self.zu + ad * D.zu))
        mu = (ga / g) * (ga / g) * (ga / self.p)

        def target(x, dx, z, dz):
            v = (x + ap_ * dx) * (z + ad_ * dz)
            tmin, tmax = mu * gamma, mu / gamma
            return np.where(v < tmin, tmin - v, np.where(v > tmax, tmax - v, 0.0))
        xzl = target(self.xl, D.xl, self.zl, D.zl)
        xzu = target(self.xu, D.xu, self.zu, D.zu)
        self._newton(Dc, z_m, z_n, z_n, z_n, xzl, xzu)
        Dc.x += D.x; Dc.xl += D.xl; Dc.xu += D.xu; Dc.y += D.y; Dc.zl += D.zl; Dc.zu += D.zu
