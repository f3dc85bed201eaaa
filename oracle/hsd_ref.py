"""CPU restatement of Tulip's homogeneous self-dual IPM driver (TEST INFRASTRUCTURE).

Restates, line by line, the *caller* of the KKT hot path so that a KKT backend can be
exercised exactly the way the reference exercises it:

* ``IPMData`` / ``standard_form``  <- src/IPM/ipmdata.jl:14-56, 64-173
* ``HSDRef.compute_residuals``     <- src/IPM/HSD/HSD.jl:77-128
* ``HSDRef.update_solver_status``  <- src/IPM/HSD/HSD.jl:136-196
* ``HSDRef.optimize``              <- src/IPM/HSD/HSD.jl:203-350
* ``HSDRef.compute_step``          <- src/IPM/HSD/step.jl:10-151
* ``HSDRef.solve_newton_system``   <- src/IPM/HSD/step.jl:198-266
* ``max_step_length``              <- src/IPM/HSD/step.jl:274-306
* ``HSDRef.compute_higher_corrector`` <- src/IPM/HSD/step.jl:325-401
* options                          <- src/IPM/options.jl:1-25

The KKT backend is any object with ``update(theta_inv, regP, regD)`` (may raise an exception
whose class name is ``PosDefException``/``ZeroPivotException``) and
``solve(dx, dy, xi_p, xi_d)``.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))


@dataclass
class IPMOptions:                      # src/IPM/options.jl:1-25
    OutputLevel: int = 0
    IterationsLimit: int = 100
    TimeLimit: float = float("inf")
    TolerancePFeas: float = SQRT_EPS
    ToleranceDFeas: float = SQRT_EPS
    ToleranceRGap: float = SQRT_EPS
    ToleranceIFeas: float = SQRT_EPS
    CorrectionLimit: int = 3
    StepDampFactor: float = 9995.0 / 10000.0
    GammaMin: float = 0.1
    CentralityOutlierThreshold: float = 0.1
    PRegMin: float = SQRT_EPS
    DRegMin: float = SQRT_EPS


@dataclass
class IPMData:                         # src/IPM/ipmdata.jl:14-56
    A: sp.csc_matrix
    b: np.ndarray
    objsense: bool
    c: np.ndarray
    c0: float
    l: np.ndarray
    u: np.ndarray
    lflag: np.ndarray = field(init=False)
    uflag: np.ndarray = field(init=False)

    def __post_init__(self):
        self.A = sp.csc_matrix(self.A, dtype=np.float64)
        self.nrow, self.ncol = self.A.shape
        self.lflag = np.isfinite(self.l)                 # ipmdata.jl:44
        self.uflag = np.isfinite(self.u)                 # ipmdata.jl:45


def standard_form(obj, obj0, objsense, rows, cols, vals, ncon, nvar, lcon, ucon, lvar, uvar):
    """src/IPM/ipmdata.jl:64-173: add slacks, flip max->min, build A (m x (n+nslack))."""
    lcon = np.asarray(lcon, float); ucon = np.asarray(ucon, float)
    b = np.zeros(ncon)
    sind, sval, lslack, uslack = [], [], [], []
    for i, (lb, ub) in enumerate(zip(lcon, ucon)):
        if lb == ub:                                       # :80-82
            b[i] = lb
        elif lb == -np.inf and ub == np.inf:               # :84-90 free row
            sind.append(i); sval.append(1.0); lslack.append(-np.inf); uslack.append(np.inf); b[i] = 0.0
        elif lb == -np.inf and np.isfinite(ub):            # :92-98  a'x + s = ub
            sind.append(i); sval.append(1.0); lslack.append(0.0); uslack.append(np.inf); b[i] = ub
        elif np.isfinite(lb) and ub == np.inf:             # :100-106 a'x - s = lb
            sind.append(i); sval.append(-1.0); lslack.append(0.0); uslack.append(np.inf); b[i] = lb
        elif np.isfinite(lb) and np.isfinite(ub):          # :108-117 range row
            sind.append(i); sval.append(1.0); lslack.append(0.0); uslack.append(ub - lb); b[i] = ub
        else:
            raise ValueError(f"Invalid bounds for row {i}: [{lb}, {ub}]")
    nslack = len(sind)
    c = np.concatenate([np.asarray(obj, float), np.zeros(nslack)])   # :129
    c0 = float(obj0)
    if not objsense:                                       # :131-135
        c = -c
        c0 = -c0
    aI = np.concatenate([np.asarray(rows, np.int64), np.asarray(sind, np.int64)])
    aJ = np.concatenate([np.asarray(cols, np.int64), nvar + np.arange(nslack, dtype=np.int64)])
    aV = np.concatenate([np.asarray(vals, float), np.asarray(sval, float)])
    A = sp.csc_matrix((aV, (aI, aJ)), shape=(ncon, nvar + nslack))
    l = np.concatenate([np.asarray(lvar, float), np.asarray(lslack, float)])
    u = np.concatenate([np.asarray(uvar, float), np.asarray(uslack, float)])
    return IPMData(A, b, bool(objsense), c, c0, l, u)


class Point:                            # src/IPM/point.jl:6-48
    def __init__(self, m, n, p):
        self.m, self.n, self.p = m, n, p
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.y = np.zeros(m); self.zl = np.zeros(n); self.zu = np.zeros(n)
        self.tau = 1.0; self.kappa = 1.0; self.mu = 1.0

    def update_mu(self):                # point.jl:45-48 (hflag = true)
        self.mu = (self.xl @ self.zl + self.xu @ self.zu + self.tau * self.kappa) / (self.p + 1)


def max_step_length_vec(x, dx):         # step.jl:274-288
    neg = dx < 0.0
    if not np.any(neg):
        return np.inf
    return float(np.min(-x[neg] / dx[neg]))


def max_step_length(pt, d):             # step.jl:295-306
    axl = max_step_length_vec(pt.xl, d.xl)
    axu = max_step_length_vec(pt.xu, d.xu)
    azl = max_step_length_vec(pt.zl, d.zl)
    azu = max_step_length_vec(pt.zu, d.zu)
    at = (-pt.tau / d.tau) if d.tau < 0.0 else 1.0
    ak = (-pt.kappa / d.kappa) if d.kappa < 0.0 else 1.0
    return min(1.0, axl, axu, azl, azu, at, ak)


def _is_factor_failure(err):
    return type(err).__name__ in ("PosDefException", "ZeroPivotException")


class HSDRef:
    """src/IPM/HSD/HSD.jl:6-65 + methods.  ``kkt`` is the plug-in backend."""

    def __init__(self, dat: IPMData, kkt, params: IPMOptions | None = None):
        self.dat = dat
        self.kkt = kkt
        self.params = params or IPMOptions()
        m, n = dat.nrow, dat.ncol
        self.p = int(dat.lflag.sum() + dat.uflag.sum())          # HSD.jl:39
        self.pt = Point(m, n, self.p)
        self.rp = np.zeros(m); self.rl = np.zeros(n); self.ru = np.zeros(n); self.rd = np.zeros(n)
        self.rg = 0.0
        self.regP = np.ones(n); self.regD = np.ones(m); self.regG = 1.0   # HSD.jl:50-52
        self.niter = 0
        self.status = "Trm_Unknown"
        self.primal_status = "Sln_Unknown"; self.dual_status = "Sln_Unknown"
        self.primal_objective = np.inf; self.dual_objective = -np.inf
        self.t_factor = 0.0; self.t_solve = 0.0; self.n_update = 0; self.n_solve = 0
        self.log = []
        # masked bounds (l .* lflag, u .* uflag) -- appear throughout step.jl
        self.lm = np.where(dat.lflag, dat.l, 0.0)
        self.um = np.where(dat.uflag, dat.u, 0.0)

    # ------------------------------------------------------------------ HSD.jl:77-128
    def compute_residuals(self):
        pt, dat = self.pt, self.dat
        self.rp = pt.tau * dat.b - dat.A @ pt.x
        self.rl = (-pt.x + pt.xl + pt.tau * self.lm) * dat.lflag
        self.ru = (-pt.x - pt.xu + pt.tau * self.um) * dat.uflag
        self.rd = pt.tau * dat.c - dat.A.T @ pt.y
        self.rd += pt.zu * dat.uflag - pt.zl * dat.lflag
        dual = dat.b @ pt.y + self.lm @ pt.zl - self.um @ pt.zu
        self.rg = pt.kappa + (dat.c @ pt.x - dual)
        inf = np.inf
        self.rp_nrm = np.linalg.norm(self.rp, inf) if len(self.rp) else 0.0
        self.rl_nrm = np.linalg.norm(self.rl, inf)
        self.ru_nrm = np.linalg.norm(self.ru, inf)
        self.rd_nrm = np.linalg.norm(self.rd, inf)
        self.rg_nrm = abs(self.rg)
        self.primal_objective = dat.c @ pt.x / pt.tau + dat.c0
        self.dual_objective = dual / pt.tau + dat.c0

    # ------------------------------------------------------------------ HSD.jl:136-196
    def update_solver_status(self):
        P = self.params
        pt, dat = self.pt, self.dat
        inf = np.inf
        self.status = "Trm_Unknown"
        nb = np.linalg.norm(dat.b, inf) if len(dat.b) else 0.0
        nl = np.linalg.norm(self.lm, inf); nu = np.linalg.norm(self.um, inf)
        nc = np.linalg.norm(dat.c, inf)
        rho_p = max(self.rp_nrm / (pt.tau * (1 + nb)),
                    self.rl_nrm / (pt.tau * (1 + nl)),
                    self.ru_nrm / (pt.tau * (1 + nu)))
        rho_d = self.rd_nrm / (pt.tau * (1 + nc))
        rho_g = abs(self.primal_objective - self.dual_objective) / (1 + abs(self.dual_objective))
        self.primal_status = "Sln_FeasiblePoint" if rho_p <= P.TolerancePFeas else "Sln_Unknown"
        self.dual_status = "Sln_FeasiblePoint" if rho_d <= P.ToleranceDFeas else "Sln_Unknown"
        if rho_p <= P.TolerancePFeas and rho_d <= P.ToleranceDFeas and rho_g <= P.ToleranceRGap:
            self.primal_status = self.dual_status = "Sln_Optimal"
            self.status = "Trm_Optimal"
            return
        Ax = dat.A @ pt.x
        lhs = max(np.linalg.norm(Ax, inf) if len(Ax) else 0.0,
                  np.linalg.norm((pt.x - pt.xl) * dat.lflag, inf),
                  np.linalg.norm((pt.x + pt.xu) * dat.uflag, inf)) * (nc / max(1.0, nb))
        if lhs < -P.ToleranceIFeas * (dat.c @ pt.x):           # HSD.jl:170-179
            self.primal_status = "Sln_InfeasibilityCertificate"
            self.status = "Trm_DualInfeasible"
            return
        delta = dat.A.T @ pt.y + pt.zl * dat.lflag - pt.zu * dat.uflag
        dualobj = dat.b @ pt.y + self.lm @ pt.zl - self.um @ pt.zu
        if np.linalg.norm(delta, inf) * max(nl, nu, nb) / max(1.0, nc) < dualobj * P.ToleranceIFeas:
            self.dual_status = "Sln_InfeasibilityCertificate"   # HSD.jl:181-192
            self.status = "Trm_PrimalInfeasible"
            return

    # ------------------------------------------------------------------ HSD.jl:203-350
    def optimize(self, callback=None):
        P = self.params
        dat, pt = self.dat, self.pt
        tstart = time.time()
        self.niter = 0
        pt.x[:] = 0.0                                            # HSD.jl:238-247
        pt.xl[:] = 1.0 * dat.lflag
        pt.xu[:] = 1.0 * dat.uflag
        pt.y[:] = 0.0
        pt.zl[:] = 1.0 * dat.lflag
        pt.zu[:] = 1.0 * dat.uflag
        pt.tau = 1.0; pt.kappa = 1.0
        pt.update_mu()
        while True:
            self.compute_residuals()                             # HSD.jl:259
            pt.update_mu()
            ttot = time.time() - tstart
            self.log.append((self.niter, self.primal_objective, self.dual_objective,
                             max(self.rp_nrm, self.ru_nrm), self.rd_nrm, self.rg_nrm, pt.mu))
            if P.OutputLevel > 0:
                eps_ = 1.0 if dat.objsense else -1.0
                print("%4d  %+14.7e  %+14.7e  %8.2e %8.2e %8.2e  %7.1e  %.2f" % (
                    self.niter, eps_ * self.primal_objective, eps_ * self.dual_objective,
                    max(self.rp_nrm, self.ru_nrm), self.rd_nrm, self.rg_nrm, pt.mu, ttot))
            self.update_solver_status()                          # HSD.jl:294
            if self.status in ("Trm_Optimal", "Trm_PrimalInfeasible", "Trm_DualInfeasible"):
                break
            elif self.niter >= P.IterationsLimit:
                self.status = "Trm_IterationLimit"; break
            elif ttot >= P.TimeLimit:
                self.status = "Trm_TimeLimit"; break
            try:
                self.compute_step()                              # HSD.jl:320
            except Exception as err:                             # HSD.jl:321-339
                nm = type(err).__name__
                if nm in ("PosDefException", "SingularException", "LinAlgError"):
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

    # ------------------------------------------------------------------ step.jl:10-151
    def compute_step(self):
        P = self.params
        dat, pt = self.dat, self.pt
        m, n, p = pt.m, pt.n, pt.p
        with np.errstate(divide="ignore", invalid="ignore"):
            thl = np.where(dat.lflag, pt.zl / pt.xl, 0.0)        # step.jl:24
            thu = np.where(dat.uflag, pt.zu / pt.xu, 0.0)        # step.jl:25
        thinv = thl + thu                                        # step.jl:26
        self.regP = np.maximum(P.PRegMin, self.regP / 10)        # step.jl:29-31
        self.regD = np.maximum(P.DRegMin, self.regD / 10)
        self.regG = max(P.PRegMin, self.regG / 10)
        nbump = 0
        while nbump <= 3:                                        # step.jl:34-51
            try:
                t0 = time.perf_counter()
                self.kkt.update(thinv, self.regP, self.regD)     # step.jl:37
                self.t_factor += time.perf_counter() - t0
                self.n_update += 1
                break
            except Exception as err:
                self.t_factor += time.perf_counter() - t0
                self.n_update += 1
                if not _is_factor_failure(err):
                    raise
                self.regD = self.regD * 100; self.regP = self.regP * 100; self.regG *= 100
                nbump += 1
        if not nbump < 3:                                        # step.jl:51 (off-by-one kept)
            from .kkt_ref import PosDefException
            raise PosDefException(0)

        self._thl, self._thu = thl, thu
        hx = np.zeros(n); hy = np.zeros(m)
        xi_ = dat.c - (thl * self.lm) - (thu * self.um)          # step.jl:61
        self._ksolve(hx, hy, dat.b, xi_)                         # step.jl:63
        # step.jl:69-76
        h0 = (self.lm @ (self.lm * thl) + self.um @ (self.um * thu)
              - (dat.c + thl * self.lm + thu * self.um) @ hx
              + dat.b @ hy + pt.kappa / pt.tau + self.regG)
        self._hx, self._hy, self._h0 = hx, hy, h0

        D = Point(m, n, p)
        self.solve_newton_system(D, hx, hy, h0,                  # step.jl:79-85
                                 self.rp, self.rl, self.ru, self.rd, self.rg,
                                 -(pt.xl * pt.zl) * dat.lflag,
                                 -(pt.xu * pt.zu) * dat.uflag,
                                 -pt.tau * pt.kappa)
        alpha = max_step_length(pt, D)                           # step.jl:88-90
        gamma = (1 - alpha) ** 2 * min(1 - alpha, P.GammaMin)
        eta = 1.0 - gamma
        self.solve_newton_system(D, hx, hy, h0,                  # step.jl:93-99
                                 eta * self.rp, eta * self.rl, eta * self.ru, eta * self.rd, eta * self.rg,
                                 (-pt.xl * pt.zl + gamma * pt.mu - D.xl * D.zl) * dat.lflag,
                                 (-pt.xu * pt.zu + gamma * pt.mu - D.xu * D.zu) * dat.uflag,
                                 -pt.tau * pt.kappa + gamma * pt.mu - D.tau * D.kappa)
        alpha = max_step_length(pt, D)                           # step.jl:100
        ncor = 0
        while ncor < P.CorrectionLimit and alpha < 0.999:        # step.jl:104-136
            alpha_ = alpha
            ncor += 1
            Dc = Point(m, n, p)
            alpha_c = self.compute_higher_corrector(Dc, gamma, hx, hy, h0, D, alpha_,
                                                    P.CentralityOutlierThreshold)
            if alpha_c > alpha_:
                D = Dc
                alpha = alpha_c
            if alpha_c < 1.1 * alpha_:
                break
        alpha *= P.StepDampFactor                                # step.jl:139-148
        pt.x += alpha * D.x; pt.xl += alpha * D.xl; pt.xu += alpha * D.xu
        pt.y += alpha * D.y; pt.zl += alpha * D.zl; pt.zu += alpha * D.zu
        pt.tau += alpha * D.tau; pt.kappa += alpha * D.kappa
        pt.update_mu()
        self.last_alpha = alpha
        self.last_ncor = ncor

    def _ksolve(self, dx, dy, xi_p, xi_d):
        t0 = time.perf_counter()
        self.kkt.solve(dx, dy, xi_p, xi_d)
        self.t_solve += time.perf_counter() - t0
        self.n_solve += 1

    # ------------------------------------------------------------------ step.jl:198-266
    def solve_newton_system(self, D, hx, hy, h0, xi_p, xi_l, xi_u, xi_d, xi_g, xi_xzl, xi_xzu, xi_tk):
        pt, dat = self.pt, self.dat
        lf, uf = dat.lflag, dat.uflag
        with np.errstate(divide="ignore", invalid="ignore"):
            ixl = np.where(lf, 1.0 / pt.xl, 0.0)
            ixu = np.where(uf, 1.0 / pt.xu, 0.0)
        # step.jl:210-213
        xi_d_ = xi_d + (-((xi_xzl + pt.zl * xi_l) * ixl) * lf + ((xi_xzu - pt.zu * xi_u) * ixu) * uf)
        self._ksolve(D.x, D.y, xi_p, xi_d_)                      # step.jl:214
        # step.jl:218-223
        xi_g_ = (xi_g + xi_tk / pt.tau
                 - ((xi_xzl * ixl) * lf) @ self.lm
                 + ((xi_xzu * ixu) * uf) @ self.um
                 - (((pt.zl * ixl) * xi_l) * lf) @ self.lm
                 - (((pt.zu * ixu) * xi_u) * uf) @ self.um)
        # step.jl:225-232
        D.tau = (xi_g_ + (dat.c + (pt.zl * ixl) * self.lm + (pt.zu * ixu) * self.um) @ D.x
                 - dat.b @ D.y) / h0
        D.x += D.tau * hx                                        # step.jl:236-237
        D.y += D.tau * hy
        D.xl = (-xi_l + D.x - D.tau * self.lm) * lf              # step.jl:240-245
        D.xu = (xi_u - D.x + D.tau * self.um) * uf
        D.zl = ((xi_xzl - pt.zl * D.xl) * ixl) * lf              # step.jl:248-249
        D.zu = ((xi_xzu - pt.zu * D.xu) * ixu) * uf
        D.kappa = (xi_tk - pt.kappa * D.tau) / pt.tau            # step.jl:252

    # ------------------------------------------------------------------ step.jl:325-401
    def compute_higher_corrector(self, Dc, gamma, hx, hy, h0, D, alpha, beta):
        pt, dat = self.pt, self.dat
        lf, uf = dat.lflag, dat.uflag
        a_ = min(1.0, 2.0 * alpha)                               # step.jl:335
        vl = ((pt.xl + a_ * D.xl) * (pt.zl + a_ * D.zl)) * lf    # step.jl:338-340
        vu = ((pt.xu + a_ * D.xu) * (pt.zu + a_ * D.zu)) * uf
        vt = (pt.tau + a_ * D.tau) * (pt.kappa + a_ * D.kappa)
        mu_l = beta * pt.mu * gamma                              # step.jl:343-344
        mu_u = gamma * pt.mu / beta

        def target(v, flag):                                     # step.jl:345-364
            out = np.where(v < mu_l, mu_l - v, np.where(v > mu_u, mu_u - v, 0.0))
            return np.where(flag, out, v)
        vl = target(vl, lf)
        vu = target(vu, uf)
        if vt < mu_l:                                            # step.jl:365-371
            vt = mu_l - vt
        elif vt > mu_u:
            vt = mu_u - vt
        else:
            vt = 0.0
        delta = (vl.sum() + vu.sum() + vt) / (pt.p + 1)          # step.jl:374-377
        vl = vl - delta
        vu = vu - delta
        vt -= delta
        m, n = pt.m, pt.n
        self.solve_newton_system(Dc, hx, hy, h0,                 # step.jl:380-386
                                 np.zeros(m), np.zeros(n), np.zeros(n), np.zeros(n), 0.0,
                                 vl, vu, vt)
        Dc.x += D.x; Dc.xl += D.xl; Dc.xu += D.xu; Dc.y += D.y   # step.jl:389-396
        Dc.zl += D.zl; Dc.zu += D.zu; Dc.tau += D.tau; Dc.kappa += D.kappa
        return max_step_length(pt, Dc)                           # step.jl:399-400
