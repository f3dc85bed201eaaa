// Device-resident main loop of Tulip's homogeneous self-dual IPM around the KKT backend.
//
// Mirrors /root/reference/src/IPM/HSD/HSD.jl:203-350 (ipm_optimize!: residuals, status tests, stopping criteria) and
// src/IPM/HSD/step.jl:10-151 (compute_step!: theta, regularisation schedule + retry loop, the 3-6 KKT solves, step length,
// Mehrotra + higher-order correctors) with every vector resident in HBM: the fused kernels of kernels_ipm.cu write the
// solver's own theta / regP / regD / rhs buffers, update! / solve! are the same (graph-replayed) device sequences the host
// API uses, and the host only reads back the scalar block (one 312-byte D2H per decision point) to take the reference's
// control-flow decisions (status tests, regularisation bumps, `alpha < 0.999` corrector loop).
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "ipm.cuh"
#include "solver_internal.hpp"

using namespace tlp;

struct tlpb200_ipm {
    IpmDev d{};
    IpmDir D[2]{};
    IpmScalars P{};
    double* h_sc = nullptr;          // pinned mirror of the device scalars
    double nb = 0, nl = 0, nu = 0, nc = 0;   // infinity norms of b, l (masked), u (masked), c   (HSD.jl:141-146)
    int32_t niter = 0;
    int32_t status = TLPB200_TRM_UNKNOWN;
    int64_t n_update = 0, n_solve = 0;
    double ms_update = 0, ms_solve = 0;       // CUDA-event time of the KKT calls ("Factorization" / "KKT" timer sections)
    std::vector<cudaEvent_t> ev;              // event pairs of the current iteration: [2k] start, [2k+1] stop
    std::vector<int> ev_kind;                 // 0 update!, 1 solve!
    size_t ev_used = 0;
    std::vector<double> log;                  // rows of 8: iter, pobj, dobj, pfeas, dfeas, gfeas, mu, tau
    bool started = false;
};

namespace {

struct CudaErr {
    cudaError_t e;
    const char* what;
};
#define CKI(call)                                          \
    do {                                                   \
        cudaError_t _e = (call);                           \
        if (_e != cudaSuccess) throw CudaErr{_e, #call};   \
    } while (0)

double* dvec(tlpb200_solver* s, size_t n) { return (double*)tlp_internal::device_alloc(s, std::max<size_t>(n, 1) * sizeof(double)); }

void mark(tlpb200_solver* s, tlpb200_ipm* ip, int kind, bool start) {
    if (start) {
        if (ip->ev_used + 2 > ip->ev.size()) {
            for (int k = 0; k < 2; ++k) { cudaEvent_t e; CKI(cudaEventCreate(&e)); ip->ev.push_back(e); }
            ip->ev_kind.resize(ip->ev.size() / 2);
        }
        ip->ev_kind[ip->ev_used / 2] = kind;
        CKI(cudaEventRecord(ip->ev[ip->ev_used], s->stream));
    } else {
        CKI(cudaEventRecord(ip->ev[ip->ev_used + 1], s->stream));
        ip->ev_used += 2;
    }
}

void collect_times(tlpb200_ipm* ip) {      // after a stream synchronisation
    for (size_t i = 0; i + 1 < ip->ev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ip->ev[i], ip->ev[i + 1]) == cudaSuccess) (ip->ev_kind[i / 2] == 0 ? ip->ms_update : ip->ms_solve) += ms;
    }
    ip->ev_used = 0;
}

void read_scalars(tlpb200_solver* s, tlpb200_ipm* ip) {
    CKI(cudaMemcpyAsync(ip->h_sc, ip->d.sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CKI(cudaStreamSynchronize(s->stream));
}

void solve(tlpb200_solver* s, tlpb200_ipm* ip) {
    mark(s, ip, 1, true);
    tlp_internal::run_solve(s);
    mark(s, ip, 1, false);
    ip->n_solve++;
}

int slot_of(int k) { return SC_DIR0 + k * DS_STRIDE; }

// step.jl:10-151.  Returns TLPB200_OK, or TLPB200_NOT_POSDEF when the factorisation could not be saved (-> Trm_NumericalProblem).
int compute_step(tlpb200_solver* s, tlpb200_ipm* ip, const tlpb200_hsd_options* o) {
    const int64_t n = s->n, m = s->m;
    cudaStream_t st = s->stream;
    IpmScalars P = ip->P;
    ipm_launch_theta(ip->d, n, m, P, s->d_theta, s->d_regP, s->d_regD, s->d_xid, s->d_xip, st);      // step.jl:24-31, :55-61
    int nbump = 0;
    while (nbump <= 3) {                                                                               // step.jl:34-51
        mark(s, ip, 0, true);
        tlp_internal::run_update(s);
        mark(s, ip, 0, false);
        ip->n_update++;
        int64_t bad = -1;
        const int rc = tlp_internal::finish_update(s, &bad);
        if (rc == TLPB200_OK) break;
        if (rc != TLPB200_NOT_POSDEF) return rc;
        ipm_launch_bump(ip->d, n, m, P, s->d_regP, s->d_regD, st);
        nbump++;
    }
    if (!(nbump < 3)) return tlp_internal::set_error(s, TLPB200_NOT_POSDEF, "factorisation could not be saved by regularisation (step.jl:51)");
    solve(s, ip);                                                                                      // step.jl:63
    ipm_launch_h(ip->d, n, m, P, s->d_dx, s->d_dy, st);                                                // step.jl:69-76
    int cur = 0;
    // affine-scaling direction (step.jl:79-85), then gamma / eta (step.jl:88-90, on the device)
    ipm_launch_newton_rhs(ip->d, n, m, P, 0, ip->D[cur], slot_of(cur), s->d_xid, s->d_xip, st);
    solve(s, ip);
    ipm_launch_newton_recover(ip->d, n, m, P, 0, ip->D[cur], slot_of(cur), ip->D[cur], slot_of(cur), s->d_dx, s->d_dy, st);
    // Mehrotra corrector (step.jl:93-99)
    ipm_launch_newton_rhs(ip->d, n, m, P, 1, ip->D[cur], slot_of(cur), s->d_xid, s->d_xip, st);
    solve(s, ip);
    ipm_launch_newton_recover(ip->d, n, m, P, 1, ip->D[cur], slot_of(cur), ip->D[cur], slot_of(cur), s->d_dx, s->d_dy, st);
    read_scalars(s, ip);
    double alpha = ip->h_sc[slot_of(cur) + DS_ALPHA];
    int ncor = 0;
    while (ncor < o->correction_limit && alpha < 0.999) {                                              // step.jl:104-136
        const double a_prev = alpha;
        ncor++;
        const int nxt = 1 - cur;
        ipm_launch_targets(ip->d, n, P, ip->D[cur], slot_of(cur), st);                                // step.jl:338-377
        ipm_launch_newton_rhs(ip->d, n, m, P, 2, ip->D[nxt], slot_of(nxt), s->d_xid, s->d_xip, st);
        solve(s, ip);                                                                                  // step.jl:380
        ipm_launch_newton_recover(ip->d, n, m, P, 2, ip->D[nxt], slot_of(nxt), ip->D[cur], slot_of(cur), s->d_dx, s->d_dy, st);
        read_scalars(s, ip);
        const double a_c = ip->h_sc[slot_of(nxt) + DS_ALPHA];
        if (a_c > a_prev) { cur = nxt; alpha = a_c; }
        if (a_c < 1.1 * a_prev) break;
    }
    ipm_launch_step(ip->d, n, m, P, ip->D[cur], slot_of(cur), st);                                     // step.jl:139-148
    CKI(cudaMemcpyAsync(s->h_info, s->ctx.info, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CKI(cudaStreamSynchronize(st));
    collect_times(ip);
    return tlp_internal::check_timeouts(s, "hsd step");
}

// HSD.jl:136-196 with the norms computed on the device
int32_t solver_status(const tlpb200_ipm* ip, const tlpb200_hsd_options* o) {
    const double* sc = ip->h_sc;
    const double tau = sc[SC_TAU];
    const double rho_p = std::fmax(sc[SC_RP_NRM] / (tau * (1 + ip->nb)), std::fmax(sc[SC_RL_NRM] / (tau * (1 + ip->nl)), sc[SC_RU_NRM] / (tau * (1 + ip->nu))));
    const double rho_d = sc[SC_RD_NRM] / (tau * (1 + ip->nc));
    const double rho_g = std::fabs(sc[SC_POBJ] - sc[SC_DOBJ]) / (1 + std::fabs(sc[SC_DOBJ]));
    if (rho_p <= o->tol_pfeas && rho_d <= o->tol_dfeas && rho_g <= o->tol_rgap) return TLPB200_TRM_OPTIMAL;
    const double lhs = std::fmax(sc[SC_AX_NRM], std::fmax(sc[SC_XXL_NRM], sc[SC_XXU_NRM])) * (ip->nc / std::fmax(1.0, ip->nb));
    if (lhs < -o->tol_ifeas * sc[SC_CX]) return TLPB200_TRM_DUAL_INFEASIBLE;
    if (sc[SC_DELTA_NRM] * std::fmax(ip->nl, std::fmax(ip->nu, ip->nb)) / std::fmax(1.0, ip->nc) < sc[SC_DUAL] * o->tol_ifeas)
        return TLPB200_TRM_PRIMAL_INFEASIBLE;
    return TLPB200_TRM_UNKNOWN;
}

void fill_info(const tlpb200_ipm* ip, tlpb200_hsd_info* out) {
    if (!out) return;
    const double* sc = ip->h_sc;
    std::memset(out, 0, sizeof *out);
    out->status = ip->status;
    out->niter = ip->niter;
    out->pobj = sc[SC_POBJ]; out->dobj = sc[SC_DOBJ];
    out->rp_nrm = sc[SC_RP_NRM]; out->rl_nrm = sc[SC_RL_NRM]; out->ru_nrm = sc[SC_RU_NRM]; out->rd_nrm = sc[SC_RD_NRM];
    out->rg_nrm = std::fabs(sc[SC_RG]);
    out->mu = sc[SC_MU]; out->tau = sc[SC_TAU]; out->kappa = sc[SC_KAPPA];
    out->n_update = ip->n_update; out->n_solve = ip->n_solve;
    out->ms_update = ip->ms_update; out->ms_solve = ip->ms_solve;
}

}  // namespace

void tlpb200_ipm_free(tlpb200_ipm* ip) {
    if (!ip) return;
    if (ip->h_sc) cudaFreeHost(ip->h_sc);
    for (auto& e : ip->ev) cudaEventDestroy(e);
    delete ip;
}

extern "C" {

#define REQUIRE_IPM(s)                                                                                             \
    if (!(s)) return TLPB200_BAD_ARG;                                                                              \
    if (!(s)->on_device) return tlp_internal::set_error((s), TLPB200_CUDA, "solver has no device state; there is no CPU fallback"); \
    if (!(s)->ipm) return tlp_internal::set_error((s), TLPB200_BAD_ARG, "tlpb200_hsd_create has not been called on this solver")

void tlpb200_hsd_default_options(tlpb200_hsd_options* o) {      // src/IPM/options.jl:1-25
    if (!o) return;
    const double sqrt_eps = std::sqrt(2.220446049250313e-16);
    o->iterations_limit = 100;
    o->correction_limit = 3;
    o->time_limit = INFINITY;
    o->tol_pfeas = o->tol_dfeas = o->tol_rgap = o->tol_ifeas = sqrt_eps;
    o->step_damp = 9995.0 / 10000.0;
    o->gamma_min = 0.1;
    o->centrality_outlier = 0.1;
    o->preg_min = o->dreg_min = sqrt_eps;
}

int tlpb200_hsd_create(tlpb200_solver* s, const double* b, const double* c, const double* l, const double* u, double c0) {
    if (!s) return TLPB200_BAD_ARG;
    if (!s->on_device) return tlp_internal::set_error(s, TLPB200_CUDA, "solver has no device state; there is no CPU fallback");
    if (!b || !c || !l || !u) return tlp_internal::set_error(s, TLPB200_BAD_ARG, "tlpb200_hsd_create: null vector");
    if (s->ipm) return tlp_internal::set_error(s, TLPB200_BAD_ARG, "tlpb200_hsd_create: already created");
    tlpb200_ipm* ip = nullptr;
    try {
        CKI(cudaSetDevice(s->device));
        ip = new tlpb200_ipm();
        const size_t n = (size_t)s->n, m = (size_t)s->m;
        std::vector<double> lm(n), um(n), lf(n), uf(n);
        double p = 0;
        for (size_t j = 0; j < n; ++j) {                     // ipmdata.jl:44-45; masked bounds as in step.jl (l .* lflag)
            lf[j] = std::isfinite(l[j]) ? 1.0 : 0.0;
            uf[j] = std::isfinite(u[j]) ? 1.0 : 0.0;
            lm[j] = lf[j] != 0.0 ? l[j] : 0.0;
            um[j] = uf[j] != 0.0 ? u[j] : 0.0;
            p += lf[j] + uf[j];
            ip->nl = std::fmax(ip->nl, std::fabs(lm[j]));
            ip->nu = std::fmax(ip->nu, std::fabs(um[j]));
            ip->nc = std::fmax(ip->nc, std::fabs(c[j]));
        }
        for (size_t i = 0; i < m; ++i) ip->nb = std::fmax(ip->nb, std::fabs(b[i]));
        auto up = [&](const double* src, size_t cnt) {
            double* q = dvec(s, cnt);
            if (cnt) CKI(cudaMemcpy(q, src, cnt * 8, cudaMemcpyHostToDevice));
            return q;
        };
        IpmDev& d = ip->d;
        d.b = up(b, m); d.c = up(c, n); d.lm = up(lm.data(), n); d.um = up(um.data(), n); d.lf = up(lf.data(), n); d.uf = up(uf.data(), n);
        d.x = dvec(s, n); d.xl = dvec(s, n); d.xu = dvec(s, n); d.y = dvec(s, m); d.zl = dvec(s, n); d.zu = dvec(s, n);
        d.rp = dvec(s, m); d.rl = dvec(s, n); d.ru = dvec(s, n); d.rd = dvec(s, n);
        d.ixl = dvec(s, n); d.ixu = dvec(s, n); d.thl = dvec(s, n); d.thu = dvec(s, n); d.cbar = dvec(s, n);
        d.hx = dvec(s, n); d.hy = dvec(s, m); d.wl = dvec(s, n); d.wu = dvec(s, n);
        d.aty_long = dvec(s, (size_t)std::max<int32_t>(s->mat.nlong, 1));
        d.sc = dvec(s, SC_COUNT);
        d.part = dvec(s, (size_t)IPM_MAXBLOCKS * IPM_NRED);
        CKI(cudaMemset(d.sc, 0, SC_COUNT * sizeof(double)));
        CKI(cudaMemset(d.part, 0, (size_t)IPM_MAXBLOCKS * IPM_NRED * sizeof(double)));
        for (auto& D : ip->D) {
            D.x = dvec(s, n); D.xl = dvec(s, n); D.xu = dvec(s, n); D.y = dvec(s, m); D.zl = dvec(s, n); D.zu = dvec(s, n);
        }
        CKI(cudaMallocHost((void**)&ip->h_sc, SC_COUNT * sizeof(double)));
        std::memset(ip->h_sc, 0, SC_COUNT * sizeof(double));
        ip->P.c0 = c0;
        ip->P.p = p;
        s->ipm = ip;
        return TLPB200_OK;
    } catch (const CudaErr& f) {
        tlpb200_ipm_free(ip);
        char buf[256];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s", (int)f.e, cudaGetErrorString(f.e), f.what);
        return tlp_internal::set_error(s, f.e == cudaErrorMemoryAllocation ? TLPB200_OOM : TLPB200_CUDA, buf);
    } catch (const std::exception& e) {
        tlpb200_ipm_free(ip);
        return tlp_internal::set_error(s, TLPB200_INTERNAL, e.what());
    }
}

// One pass of the main-loop body (HSD.jl:256-320): residuals, status tests, stopping criteria, then -- unless the run is
// over -- compute_step!.  info->status == TLPB200_TRM_UNKNOWN means "iterate again".  The first call starts from the
// reference's start point (HSD.jl:238-247); tlpb200_hsd_reset restarts.
int tlpb200_hsd_iterate(tlpb200_solver* s, const tlpb200_hsd_options* opt, tlpb200_hsd_info* info) {
    REQUIRE_IPM(s);
    tlpb200_ipm* ip = s->ipm;
    tlpb200_hsd_options o;
    if (opt) o = *opt; else tlpb200_hsd_default_options(&o);
    try {
        CKI(cudaSetDevice(s->device));
        ip->P.preg_min = o.preg_min; ip->P.dreg_min = o.dreg_min; ip->P.gamma_min = o.gamma_min;
        ip->P.beta = o.centrality_outlier; ip->P.step_damp = o.step_damp;
        if (!ip->started) {
            ipm_launch_start(ip->d, s->n, s->m, s->d_regP, s->d_regD, s->stream);
            ip->started = true;
            ip->niter = 0; ip->n_update = ip->n_solve = 0; ip->ms_update = ip->ms_solve = 0;
            ip->log.clear();
        }
        ipm_launch_residuals(ip->d, s->mat, ip->P, s->stream);                    // HSD.jl:259 + update_mu!
        read_scalars(s, ip);
        const double* sc = ip->h_sc;
        const double row[8] = {(double)ip->niter, sc[SC_POBJ], sc[SC_DOBJ], std::fmax(sc[SC_RP_NRM], sc[SC_RU_NRM]), sc[SC_RD_NRM],
                               std::fabs(sc[SC_RG]), sc[SC_MU], sc[SC_TAU]};
        ip->log.insert(ip->log.end(), row, row + 8);
        ip->status = solver_status(ip, &o);                                       // HSD.jl:294
        if (ip->status == TLPB200_TRM_UNKNOWN && ip->niter >= o.iterations_limit) ip->status = TLPB200_TRM_ITERATION_LIMIT;
        if (ip->status == TLPB200_TRM_UNKNOWN) {
            const int rc = compute_step(s, ip, &o);                               // HSD.jl:320
            if (rc == TLPB200_NOT_POSDEF) ip->status = TLPB200_TRM_NUMERICAL_PROBLEM;   // HSD.jl:321-326
            else if (rc == TLPB200_OOM) ip->status = TLPB200_TRM_MEMORY_LIMIT;          // HSD.jl:327
            else if (rc != TLPB200_OK) { fill_info(ip, info); return rc; }
            else ip->niter++;
        }
        fill_info(ip, info);
        return TLPB200_OK;
    } catch (const CudaErr& f) {
        char buf[256];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s", (int)f.e, cudaGetErrorString(f.e), f.what);
        cudaGetLastError();
        return tlp_internal::set_error(s, f.e == cudaErrorMemoryAllocation ? TLPB200_OOM : TLPB200_CUDA, buf);
    } catch (const std::exception& e) {
        return tlp_internal::set_error(s, TLPB200_INTERNAL, s->err.empty() ? e.what() : s->err);
    }
}

// ipm_optimize! (HSD.jl:203-350): iterate until a status is reached or a limit hits
int tlpb200_hsd_optimize(tlpb200_solver* s, const tlpb200_hsd_options* opt, tlpb200_hsd_info* info) {
    REQUIRE_IPM(s);
    tlpb200_hsd_options o;
    if (opt) o = *opt; else tlpb200_hsd_default_options(&o);
    s->ipm->started = false;
    const auto t0 = std::chrono::steady_clock::now();
    tlpb200_hsd_info cur{};
    while (true) {
        const int rc = tlpb200_hsd_iterate(s, &o, &cur);
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        cur.seconds_total = el;
        if (info) *info = cur;
        if (rc != TLPB200_OK) return rc;
        if (cur.status != TLPB200_TRM_UNKNOWN) break;
        if (el >= o.time_limit) { s->ipm->status = TLPB200_TRM_TIME_LIMIT; cur.status = TLPB200_TRM_TIME_LIMIT; if (info) *info = cur; break; }
    }
    return TLPB200_OK;
}

int tlpb200_hsd_reset(tlpb200_solver* s) {
    REQUIRE_IPM(s);
    s->ipm->started = false;
    return TLPB200_OK;
}

int tlpb200_hsd_get_point(tlpb200_solver* s, double* x, double* xl, double* xu, double* y, double* zl, double* zu, double* tau_kappa) {
    REQUIRE_IPM(s);
    const tlpb200_ipm* ip = s->ipm;
    const size_t n = (size_t)s->n, m = (size_t)s->m;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaError_t e = cudaSuccess;
    auto dl = [&](double* dst, const double* src, size_t cnt) { if (dst && cnt && e == cudaSuccess) e = cudaMemcpy(dst, src, cnt * 8, cudaMemcpyDeviceToHost); };
    dl(x, ip->d.x, n); dl(xl, ip->d.xl, n); dl(xu, ip->d.xu, n); dl(y, ip->d.y, m); dl(zl, ip->d.zl, n); dl(zu, ip->d.zu, n);
    if (tau_kappa && e == cudaSuccess) {
        double sc[SC_COUNT];
        e = cudaMemcpy(sc, ip->d.sc, sizeof sc, cudaMemcpyDeviceToHost);
        tau_kappa[0] = sc[SC_TAU]; tau_kappa[1] = sc[SC_KAPPA];
    }
    if (e != cudaSuccess) return tlp_internal::set_error(s, TLPB200_CUDA, cudaGetErrorString(e));
    return TLPB200_OK;
}

/* per-iteration log of the run so far: rows of 8 doubles {iter, pobj, dobj, pfeas, dfeas, gfeas, mu, tau} (the columns of the
   reference's iteration line, HSD.jl:266-287); *rows receives the count, out may be NULL */
int tlpb200_hsd_get_log(tlpb200_solver* s, double* out, int64_t* rows) {
    REQUIRE_IPM(s);
    const tlpb200_ipm* ip = s->ipm;
    if (rows) *rows = (int64_t)(ip->log.size() / 8);
    if (out && !ip->log.empty()) std::memcpy(out, ip->log.data(), ip->log.size() * sizeof(double));
    return TLPB200_OK;
}

}  // extern "C"
