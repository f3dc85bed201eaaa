// Device-resident iteration of Tulip's homogeneous self-dual IPM around the KKT hot path.
//
// The reference runs these steps as ~30-40 host vector passes per Newton solve with fresh allocations
// (/root/reference/src/IPM/HSD/step.jl:24-31 theta + regularisation schedule, :61 xi_ for the first solve, :69-76 h0,
// :210-213 xi_d_, :218-252 recoveries, :274-306 max_step_length, :338-377 corrector targets; HSD.jl:77-128 residuals,
// :136-196 the quantities of the status tests).  Here each of them is ONE fused elementwise kernel over max(n, m) indices
// with its dot products / max-norms / step-length minima reduced in the same pass (deterministic two-stage reduction:
// per-block partials, then a one-block finish kernel that also does the scalar algebra), so that between update! and the
// last solve! of an iteration the vectors never leave HBM and the host reads back a handful of scalars.
//
// Every kernel is HBM-bound elementwise work: 8 B per vector element read or written, coalesced; the sparse products
// A x / A'y inside k_ipm_residuals read A once in CSR and once in CSC (12 B per non-zero each).
#include <cuda_runtime.h>

#include <cstdint>

#include "ipm.cuh"

namespace tlp {

namespace {

constexpr int IPM_THREADS = 256;

enum RedOp { RSUM = 0, RMAX = 1, RMIN = 2 };

__device__ __forceinline__ double red_op(double a, double b, int op) {
    return op == RSUM ? a + b : (op == RMAX ? fmax(a, b) : fmin(a, b));
}
__device__ __forceinline__ double red_identity(int op) { return op == RSUM ? 0.0 : (op == RMAX ? 0.0 : 1.0e300); }

// block-reduce NR per-thread values (ops[k] says how) and store them to part[blockIdx.x][k]
template <int NR>
__device__ __forceinline__ void block_partials(double (&v)[NR], const int (&ops)[NR], double* __restrict__ part) {
    __shared__ double sh[IPM_THREADS / 32][IPM_NRED];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = red_op(x, __shfl_xor_sync(0xffffffffu, x, o), ops[k]);
        if (lane == 0) sh[w][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NR) {
        double x = sh[0][threadIdx.x];
        for (int q = 1; q < IPM_THREADS / 32; ++q) x = red_op(x, sh[q][threadIdx.x], ops[threadIdx.x]);
        part[(size_t)blockIdx.x * IPM_NRED + threadIdx.x] = x;
    }
}

// |x| with NaN made visible to a max-norm (fmax would drop it)
__device__ __forceinline__ double absn(double x) { return x != x ? 1.0e300 : fabs(x); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// residuals + the quantities of the status tests   (HSD.jl:77-128, :136-196; point.jl:45-48)
// slots: 0 max|rl| 1 max|ru| 2 max|rd| 3 c'x 4 lm'zl 5 um'zu 6 xl'zl + xu'zu 7 max|(x-xl) lf| 8 max|(x+xu) uf|
//        9 max|A'y + zl lf - zu uf| 10 max|rp| 11 b'y 12 max|A x|
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_residuals(IpmDev d, DevMat A) {
    const double tau = d.sc[SC_TAU];
    double v[13];
    const int ops[13] = {RMAX, RMAX, RMAX, RSUM, RSUM, RSUM, RSUM, RMAX, RMAX, RMAX, RMAX, RSUM, RMAX};
#pragma unroll
    for (int k = 0; k < 13; ++k) v[k] = 0.0;
    const int64_t mx = A.n > A.m ? A.n : A.m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < A.n) {
            double aty = 0.0;
            const int64_t cb = A.colptr[i], ce = A.colptr[i + 1];
            if (ce - cb <= LONG_COL) {
                for (int64_t p = cb; p < ce; ++p) aty += A.val[p] * d.y[A.rowidx[p]];
            } else {          // long column: A'y was formed by k_ipm_aty_long (one CTA per column); find its slot
                int lo = 0, hi = A.nlong - 1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.long_cols[mid] < (int32_t)i) lo = mid + 1; else hi = mid; }
                aty = d.aty_long[lo];
            }
            const double lf = d.lf[i], uf = d.uf[i], x = d.x[i], xl = d.xl[i], xu = d.xu[i], zl = d.zl[i], zu = d.zu[i];
            const double lm = d.lm[i], um = d.um[i], c = d.c[i];
            const double rl = lf != 0.0 ? (-x + xl + tau * lm) : 0.0;
            const double ru = uf != 0.0 ? (-x - xu + tau * um) : 0.0;
            const double rd = tau * c - aty + (uf != 0.0 ? zu : 0.0) - (lf != 0.0 ? zl : 0.0);
            d.rl[i] = rl; d.ru[i] = ru; d.rd[i] = rd;
            v[0] = fmax(v[0], absn(rl)); v[1] = fmax(v[1], absn(ru)); v[2] = fmax(v[2], absn(rd));
            v[3] += c * x; v[4] += lm * zl; v[5] += um * zu; v[6] += xl * zl + xu * zu;
            v[7] = fmax(v[7], lf != 0.0 ? absn(x - xl) : 0.0);
            v[8] = fmax(v[8], uf != 0.0 ? absn(x + xu) : 0.0);
            v[9] = fmax(v[9], absn(aty + (lf != 0.0 ? zl : 0.0) - (uf != 0.0 ? zu : 0.0)));
        }
        if (i < A.m) {
            double ax = 0.0;
            for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) ax += A.rval[p] * d.x[A.colidx[p]];
            const double b = d.b[i];
            const double rp = tau * b - ax;
            d.rp[i] = rp;
            v[10] = fmax(v[10], absn(rp)); v[11] += b * d.y[i]; v[12] = fmax(v[12], absn(ax));
        }
    }
    block_partials<13>(v, ops, d.part);
}

// A'y of the long columns, one CTA per column (feeds k_ipm_residuals)
__global__ void __launch_bounds__(256) k_ipm_aty_long(IpmDev d, DevMat A) {
    __shared__ double red[8];
    const int64_t j = A.long_cols[blockIdx.x];
    double v = 0.0;
    for (int64_t p = A.colptr[j] + threadIdx.x; p < A.colptr[j + 1]; p += blockDim.x) v += A.val[p] * d.y[A.rowidx[p]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        d.aty_long[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// theta, regularisation schedule, rhs of the first solve, cbar   (step.jl:24-31, :55-61)
// slots: 0 lm'(lm thl) + um'(um thu)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_theta(IpmDev d, int64_t n, int64_t m, double preg_min, double dreg_min,
                                                          double* __restrict__ theta, double* __restrict__ regP,
                                                          double* __restrict__ regD, double* __restrict__ xi_d, double* __restrict__ xi_p) {
    double v[1] = {0.0};
    const int ops[1] = {RSUM};
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            const double lf = d.lf[i], uf = d.uf[i];
            const double ixl = lf != 0.0 ? 1.0 / d.xl[i] : 0.0, ixu = uf != 0.0 ? 1.0 / d.xu[i] : 0.0;
            const double thl = d.zl[i] * ixl, thu = d.zu[i] * ixu;
            const double lm = d.lm[i], um = d.um[i], c = d.c[i];
            d.ixl[i] = ixl; d.ixu[i] = ixu; d.thl[i] = thl; d.thu[i] = thu;
            theta[i] = thl + thu;
            regP[i] = fmax(preg_min, regP[i] / 10.0);
            xi_d[i] = c - thl * lm - thu * um;
            d.cbar[i] = c + thl * lm + thu * um;
            v[0] += lm * (lm * thl) + um * (um * thu);
        }
        if (i < m) {
            regD[i] = fmax(dreg_min, regD[i] / 10.0);
            xi_p[i] = d.b[i];
        }
    }
    block_partials<1>(v, ops, d.part);
}

// regularisation bump of the retry loop (step.jl:43-45)
__global__ void k_ipm_scale_regs(double* __restrict__ regP, int64_t n, double* __restrict__ regD, int64_t m, double f) {
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) regP[i] *= f;
        if (i < m) regD[i] *= f;
    }
}

// (hx, hy) <- solution of the first solve; slots: 0 cbar'hx  1 b'hy   (step.jl:63, :69-76)
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_h(IpmDev d, int64_t n, int64_t m, const double* __restrict__ dx,
                                                      const double* __restrict__ dy) {
    double v[2] = {0.0, 0.0};
    const int ops[2] = {RSUM, RSUM};
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) { const double h = dx[i]; d.hx[i] = h; v[0] += d.cbar[i] * h; }
        if (i < m) { const double h = dy[i]; d.hy[i] = h; v[1] += d.b[i] * h; }
    }
    block_partials<2>(v, ops, d.part);
}

// ---------------------------------------------------------------------------------------------------------------------
// right-hand side of a Newton system   (step.jl:79-99 callers, :210-223 body, :338-377 corrector targets)
//   mode 0 affine:    xi = residuals, xi_xz = -x z
//   mode 1 corrector: xi = eta residuals, xi_xz = -x z + gamma mu - dx dz of the affine direction (held in D)
//   mode 2 higher-order corrector: xi = 0, xi_xz = target(v) - delta (targets already in wl / wu)
// slots: 0 (wl ixl)'lm  1 (wu ixu)'um  2 (thl xi_l)'lm  3 (thu xi_u)'um
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_newton_rhs(IpmDev d, int64_t n, int64_t m, int mode, IpmDir D,
                                                               double* __restrict__ xi_d, double* __restrict__ xi_p) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const int ops[4] = {RSUM, RSUM, RSUM, RSUM};
    const double sigma = mode == 0 ? 1.0 : (mode == 1 ? d.sc[SC_ETA] : 0.0);
    const double gmu = d.sc[SC_GAMMA] * d.sc[SC_MU];
    const double delta = d.sc[SC_DELTA];
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            const double lf = d.lf[i], uf = d.uf[i], zl = d.zl[i], zu = d.zu[i], ixl = d.ixl[i], ixu = d.ixu[i];
            double wl, wu;
            if (mode == 0) {
                wl = lf != 0.0 ? -d.xl[i] * zl : 0.0;
                wu = uf != 0.0 ? -d.xu[i] * zu : 0.0;
            } else if (mode == 1) {
                wl = lf != 0.0 ? (-d.xl[i] * zl + gmu - D.xl[i] * D.zl[i]) : 0.0;
                wu = uf != 0.0 ? (-d.xu[i] * zu + gmu - D.xu[i] * D.zu[i]) : 0.0;
            } else {
                wl = lf != 0.0 ? d.wl[i] - delta : 0.0;
                wu = uf != 0.0 ? d.wu[i] - delta : 0.0;
            }
            d.wl[i] = wl; d.wu[i] = wu;
            const double xil = sigma * d.rl[i], xiu = sigma * d.ru[i];
            xi_d[i] = sigma * d.rd[i] - (wl + zl * xil) * ixl + (wu - zu * xiu) * ixu;
            const double lm = d.lm[i], um = d.um[i];
            v[0] += (wl * ixl) * lm; v[1] += (wu * ixu) * um; v[2] += (d.thl[i] * xil) * lm; v[3] += (d.thu[i] * xiu) * um;
        }
        if (i < m) xi_p[i] = sigma * d.rp[i];
    }
    block_partials<4>(v, ops, d.part);
}

// slots: 0 cbar'dx  1 b'dy   (step.jl:225-232)
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_dots(IpmDev d, int64_t n, int64_t m, const double* __restrict__ dx,
                                                         const double* __restrict__ dy) {
    double v[2] = {0.0, 0.0};
    const int ops[2] = {RSUM, RSUM};
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) v[0] += d.cbar[i] * dx[i];
        if (i < m) v[1] += d.b[i] * dy[i];
    }
    block_partials<2>(v, ops, d.part);
}

// recovery of the full direction + maximum step length   (step.jl:234-252, :274-306)
//   D.x = dx + dtau hx ...; mode 2 adds the previous direction P (compute_higher_corrector!, step.jl:384-391)
// slots: 0 min over {xl, xu, zl, zu} of -v/dv where dv < 0
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_newton_recover(IpmDev d, int64_t n, int64_t m, int mode, IpmDir D, IpmDir P,
                                                                   const double* __restrict__ dx, const double* __restrict__ dy) {
    double v[1] = {1.0e300};
    const int ops[1] = {RMIN};
    const double sigma = mode == 0 ? 1.0 : (mode == 1 ? d.sc[SC_ETA] : 0.0);
    const double dtau = d.sc[SC_DTAU_NEW];
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            const double lf = d.lf[i], uf = d.uf[i];
            const double xil = sigma * d.rl[i], xiu = sigma * d.ru[i];
            double x = dx[i] + dtau * d.hx[i];
            double xl = lf != 0.0 ? (-xil + x - dtau * d.lm[i]) : 0.0;
            double xu = uf != 0.0 ? (xiu - x + dtau * d.um[i]) : 0.0;
            double zl = (d.wl[i] - d.zl[i] * xl) * d.ixl[i];
            double zu = (d.wu[i] - d.zu[i] * xu) * d.ixu[i];
            if (mode == 2) { x += P.x[i]; xl += P.xl[i]; xu += P.xu[i]; zl += P.zl[i]; zu += P.zu[i]; }
            D.x[i] = x; D.xl[i] = xl; D.xu[i] = xu; D.zl[i] = zl; D.zu[i] = zu;
            if (xl < 0.0) v[0] = fmin(v[0], -d.xl[i] / xl);
            if (xu < 0.0) v[0] = fmin(v[0], -d.xu[i] / xu);
            if (zl < 0.0) v[0] = fmin(v[0], -d.zl[i] / zl);
            if (zu < 0.0) v[0] = fmin(v[0], -d.zu[i] / zu);
        }
        if (i < m) {
            double y = dy[i] + dtau * d.hy[i];
            if (mode == 2) y += P.y[i];
            D.y[i] = y;
        }
    }
    block_partials<1>(v, ops, d.part);
}

// corrector targets   (step.jl:338-377);  slots: 0 sum of the targets
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_targets(IpmDev d, int64_t n, IpmDir D) {
    double v[1] = {0.0};
    const int ops[1] = {RSUM};
    const double a_ = d.sc[SC_ALPHA_];
    const double mu_l = d.sc[SC_MU_L], mu_u = d.sc[SC_MU_U];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double tl = 0.0, tu = 0.0;
        if (d.lf[i] != 0.0) {
            const double p = (d.xl[i] + a_ * D.xl[i]) * (d.zl[i] + a_ * D.zl[i]);
            tl = p < mu_l ? mu_l - p : (p > mu_u ? mu_u - p : 0.0);
        }
        if (d.uf[i] != 0.0) {
            const double p = (d.xu[i] + a_ * D.xu[i]) * (d.zu[i] + a_ * D.zu[i]);
            tu = p < mu_l ? mu_l - p : (p > mu_u ? mu_u - p : 0.0);
        }
        d.wl[i] = tl; d.wu[i] = tu;
        v[0] += tl + tu;
    }
    block_partials<1>(v, ops, d.part);
}

// pt <- pt + alpha D   (step.jl:139-148); alpha (already damped) is read from sc[SC_STEP]
__global__ void __launch_bounds__(IPM_THREADS) k_ipm_update_point(IpmDev d, int64_t n, int64_t m, IpmDir D) {
    const double a = d.sc[SC_STEP];
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            d.x[i] += a * D.x[i]; d.xl[i] += a * D.xl[i]; d.xu[i] += a * D.xu[i];
            d.zl[i] += a * D.zl[i]; d.zu[i] += a * D.zu[i];
        }
        if (i < m) d.y[i] += a * D.y[i];
    }
}

// start point   (HSD.jl:238-247)
__global__ void k_ipm_start(IpmDev d, int64_t n, int64_t m, double* __restrict__ regP, double* __restrict__ regD) {
    const int64_t mx = n > m ? n : m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < mx; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) {
            d.x[i] = 0.0; d.xl[i] = d.lf[i]; d.xu[i] = d.uf[i]; d.zl[i] = d.lf[i]; d.zu[i] = d.uf[i];
            regP[i] = 1.0;
        }
        if (i < m) { d.y[i] = 0.0; regD[i] = 1.0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        d.sc[SC_TAU] = 1.0; d.sc[SC_KAPPA] = 1.0; d.sc[SC_REGG] = 1.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one-block finish kernel: second stage of the reductions + the scalar algebra of the phase
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_ipm_finish(IpmDev d, int phase, int nblocks, int nred, IpmOps ops, int slot_dir, int slot_prev,
                                                   IpmScalars P) {
    __shared__ double r[IPM_NRED];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w < nred) {
        const int op = ops.op[w];
        double x = red_identity(op);
        if (op == RMAX) x = 0.0;
        for (int b = lane; b < nblocks; b += 32) x = red_op(x, d.part[(size_t)b * IPM_NRED + w], op);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = red_op(x, __shfl_xor_sync(0xffffffffu, x, o), op);
        if (lane == 0) r[w] = x;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double* sc = d.sc;
    const double tau = sc[SC_TAU], kappa = sc[SC_KAPPA];
    if (phase == PH_RESIDUALS) {
        const double cx = r[3];
        const double dual = r[11] + r[4] - r[5];
        sc[SC_RL_NRM] = r[0]; sc[SC_RU_NRM] = r[1]; sc[SC_RD_NRM] = r[2]; sc[SC_RP_NRM] = r[10];
        sc[SC_CX] = cx; sc[SC_DUAL] = dual;
        sc[SC_RG] = kappa + (cx - dual);
        sc[SC_POBJ] = cx / tau + P.c0;
        sc[SC_DOBJ] = dual / tau + P.c0;
        sc[SC_MU] = (r[6] + tau * kappa) / (P.p + 1.0);
        sc[SC_XXL_NRM] = r[7]; sc[SC_XXU_NRM] = r[8]; sc[SC_DELTA_NRM] = r[9]; sc[SC_AX_NRM] = r[12];
    } else if (phase == PH_THETA) {
        sc[SC_S0] = r[0];
        sc[SC_REGG] = fmax(P.preg_min, sc[SC_REGG] / 10.0);
    } else if (phase == PH_BUMP) {
        sc[SC_REGG] *= 100.0;
    } else if (phase == PH_H0) {
        sc[SC_H0] = sc[SC_S0] - r[0] + r[1] + kappa / tau + sc[SC_REGG];
    } else if (phase == PH_RHS) {
        // xi_g_ (step.jl:218-223) without the dot products of the solve; mode in P.mode
        const double sigma = P.mode == 0 ? 1.0 : (P.mode == 1 ? sc[SC_ETA] : 0.0);
        double xi_tk;
        if (P.mode == 0) xi_tk = -tau * kappa;
        else if (P.mode == 1) xi_tk = -tau * kappa + sc[SC_GAMMA] * sc[SC_MU] - sc[slot_dir + DS_TAU] * sc[slot_dir + DS_KAPPA];
        else xi_tk = sc[SC_VT] - sc[SC_DELTA];
        sc[SC_XI_TK] = xi_tk;
        sc[SC_XI_G] = sigma * sc[SC_RG] + xi_tk / tau - r[0] + r[1] - r[2] - r[3];
    } else if (phase == PH_DTAU) {
        const double dtau = (sc[SC_XI_G] + r[0] - r[1]) / sc[SC_H0];
        sc[SC_DTAU_NEW] = dtau;
        sc[SC_DKAPPA_NEW] = (sc[SC_XI_TK] - kappa * dtau) / tau;
    } else if (phase == PH_ALPHA) {
        double dtau = sc[SC_DTAU_NEW], dkappa = sc[SC_DKAPPA_NEW];
        if (P.mode == 2) { dtau += sc[slot_prev + DS_TAU]; dkappa += sc[slot_prev + DS_KAPPA]; }
        sc[slot_dir + DS_TAU] = dtau; sc[slot_dir + DS_KAPPA] = dkappa;
        double a = fmin(1.0, r[0]);
        if (dtau < 0.0) a = fmin(a, -tau / dtau);
        if (dkappa < 0.0) a = fmin(a, -kappa / dkappa);
        sc[slot_dir + DS_ALPHA] = a;
        if (P.mode == 0) {          // step.jl:88-90
            const double g = (1.0 - a) * (1.0 - a) * fmin(1.0 - a, P.gamma_min);
            sc[SC_GAMMA] = g;
            sc[SC_ETA] = 1.0 - g;
        }
    } else if (phase == PH_TARGET_PRE) {
        // alpha_ and the target box of a higher-order corrector (step.jl:338-343); alpha of the current direction in slot_dir
        const double a_ = fmin(1.0, 2.0 * sc[slot_dir + DS_ALPHA]);
        sc[SC_ALPHA_] = a_;
        sc[SC_MU_L] = P.beta * sc[SC_MU] * sc[SC_GAMMA];
        sc[SC_MU_U] = sc[SC_GAMMA] * sc[SC_MU] / P.beta;
    } else if (phase == PH_TARGET) {
        const double a_ = sc[SC_ALPHA_], mu_l = sc[SC_MU_L], mu_u = sc[SC_MU_U];
        const double p = (tau + a_ * sc[slot_dir + DS_TAU]) * (kappa + a_ * sc[slot_dir + DS_KAPPA]);
        const double vt = p < mu_l ? mu_l - p : (p > mu_u ? mu_u - p : 0.0);
        sc[SC_VT] = vt;
        sc[SC_DELTA] = (r[0] + vt) / (P.p + 1.0);
    } else if (phase == PH_STEP) {
        const double a = sc[slot_dir + DS_ALPHA] * P.step_damp;
        sc[SC_STEP] = a;
        sc[SC_TAU] = tau + a * sc[slot_dir + DS_TAU];
        sc[SC_KAPPA] = kappa + a * sc[slot_dir + DS_KAPPA];
    }
}

// ---- launchers ----------------------------------------------------------------------------------------------------------
static inline int ipm_grid(int64_t n) {
    const int64_t b = (n + IPM_THREADS - 1) / IPM_THREADS;
    return (int)(b < 1 ? 1 : (b > IPM_MAXBLOCKS ? IPM_MAXBLOCKS : b));
}

static void finish(const IpmDev& d, int phase, int nblocks, int nred, const int* ops, int slot_dir, int slot_prev, const IpmScalars& P,
                   cudaStream_t st) {
    IpmOps o{};
    for (int k = 0; k < nred; ++k) o.op[k] = ops[k];
    k_ipm_finish<<<1, 512, 0, st>>>(d, phase, nblocks, nred, o, slot_dir, slot_prev, P);
}

void ipm_launch_start(const IpmDev& d, int64_t n, int64_t m, double* regP, double* regD, cudaStream_t st) {
    k_ipm_start<<<ipm_grid(n > m ? n : m), IPM_THREADS, 0, st>>>(d, n, m, regP, regD);
}

void ipm_launch_residuals(const IpmDev& d, const DevMat& A, const IpmScalars& P, cudaStream_t st) {
    const int g = ipm_grid(A.n > A.m ? A.n : A.m);
    if (A.nlong > 0) k_ipm_aty_long<<<A.nlong, 256, 0, st>>>(d, A);
    k_ipm_residuals<<<g, IPM_THREADS, 0, st>>>(d, A);
    const int ops[13] = {RMAX, RMAX, RMAX, RSUM, RSUM, RSUM, RSUM, RMAX, RMAX, RMAX, RMAX, RSUM, RMAX};
    finish(d, PH_RESIDUALS, g, 13, ops, 0, 0, P, st);
}

void ipm_launch_theta(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, double* theta, double* regP, double* regD,
                      double* xi_d, double* xi_p, cudaStream_t st) {
    const int g = ipm_grid(n > m ? n : m);
    k_ipm_theta<<<g, IPM_THREADS, 0, st>>>(d, n, m, P.preg_min, P.dreg_min, theta, regP, regD, xi_d, xi_p);
    const int ops[1] = {RSUM};
    finish(d, PH_THETA, g, 1, ops, 0, 0, P, st);
}

void ipm_launch_bump(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, double* regP, double* regD, cudaStream_t st) {
    k_ipm_scale_regs<<<ipm_grid(n > m ? n : m), IPM_THREADS, 0, st>>>(regP, n, regD, m, 100.0);
    finish(d, PH_BUMP, 0, 0, nullptr, 0, 0, P, st);
}

void ipm_launch_h(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, const double* dx, const double* dy, cudaStream_t st) {
    const int g = ipm_grid(n > m ? n : m);
    k_ipm_h<<<g, IPM_THREADS, 0, st>>>(d, n, m, dx, dy);
    const int ops[2] = {RSUM, RSUM};
    finish(d, PH_H0, g, 2, ops, 0, 0, P, st);
}

void ipm_launch_newton_rhs(const IpmDev& d, int64_t n, int64_t m, IpmScalars P, int mode, const IpmDir& D, int slot_dir,
                           double* xi_d, double* xi_p, cudaStream_t st) {
    const int g = ipm_grid(n > m ? n : m);
    P.mode = mode;
    k_ipm_newton_rhs<<<g, IPM_THREADS, 0, st>>>(d, n, m, mode, D, xi_d, xi_p);
    const int ops[4] = {RSUM, RSUM, RSUM, RSUM};
    finish(d, PH_RHS, g, 4, ops, slot_dir, 0, P, st);
}

void ipm_launch_newton_recover(const IpmDev& d, int64_t n, int64_t m, IpmScalars P, int mode, const IpmDir& D, int slot_dir,
                               const IpmDir& Pv, int slot_prev, const double* dx, const double* dy, cudaStream_t st) {
    const int g = ipm_grid(n > m ? n : m);
    P.mode = mode;
    k_ipm_dots<<<g, IPM_THREADS, 0, st>>>(d, n, m, dx, dy);
    const int ops2[2] = {RSUM, RSUM};
    finish(d, PH_DTAU, g, 2, ops2, slot_dir, slot_prev, P, st);
    k_ipm_newton_recover<<<g, IPM_THREADS, 0, st>>>(d, n, m, mode, D, Pv, dx, dy);
    const int ops1[1] = {RMIN};
    finish(d, PH_ALPHA, g, 1, ops1, slot_dir, slot_prev, P, st);
}

void ipm_launch_targets(const IpmDev& d, int64_t n, const IpmScalars& P, const IpmDir& D, int slot_dir, cudaStream_t st) {
    const int g = ipm_grid(n);
    finish(d, PH_TARGET_PRE, 0, 0, nullptr, slot_dir, 0, P, st);
    k_ipm_targets<<<g, IPM_THREADS, 0, st>>>(d, n, D);
    const int ops[1] = {RSUM};
    finish(d, PH_TARGET, g, 1, ops, slot_dir, 0, P, st);
}

void ipm_launch_step(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, const IpmDir& D, int slot_dir, cudaStream_t st) {
    finish(d, PH_STEP, 0, 0, nullptr, slot_dir, 0, P, st);
    k_ipm_update_point<<<ipm_grid(n > m ? n : m), IPM_THREADS, 0, st>>>(d, n, m, D);
}

}  // namespace tlp
