// Device-resident HSD iteration: shared declarations of kernels_ipm.cu (kernels) and ipm.cu (host loop + C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace tlp {

constexpr int IPM_NRED = 16;          // reduction slots per kernel (per-block partials are [blocks][IPM_NRED])
constexpr int IPM_MAXBLOCKS = 592;    // 4 CTAs of 256 threads per SM on 148 SMs; grid-stride loops beyond that

// device scalars (doubles), one array `sc`
enum IpmScalarSlot {
    SC_TAU = 0, SC_KAPPA, SC_MU, SC_REGG, SC_S0, SC_H0, SC_GAMMA, SC_ETA, SC_DELTA, SC_VT, SC_ALPHA_, SC_MU_L, SC_MU_U,
    SC_XI_TK, SC_XI_G, SC_DTAU_NEW, SC_DKAPPA_NEW, SC_STEP,
    // residual / status block: read back by the host once per iteration (HSD.jl:136-196)
    SC_RP_NRM, SC_RL_NRM, SC_RU_NRM, SC_RD_NRM, SC_RG, SC_POBJ, SC_DOBJ, SC_CX, SC_DUAL, SC_AX_NRM, SC_XXL_NRM, SC_XXU_NRM,
    SC_DELTA_NRM,
    SC_DIR0,            // direction-scalar blocks: {tau, kappa, alpha} x 2 directions
    SC_COUNT = SC_DIR0 + 8
};
enum IpmDirSlot { DS_TAU = 0, DS_KAPPA = 1, DS_ALPHA = 2, DS_STRIDE = 4 };

enum IpmPhase { PH_RESIDUALS = 0, PH_THETA, PH_BUMP, PH_H0, PH_RHS, PH_DTAU, PH_ALPHA, PH_TARGET_PRE, PH_TARGET, PH_STEP };

struct IpmOps {
    int op[IPM_NRED];
};

// host-side constants passed by value to the finish kernel
struct IpmScalars {
    double c0, p;                 // objective offset, number of finite bounds (HSD.jl:39)
    double preg_min, dreg_min;    // options.jl:21-22
    double gamma_min, beta, step_damp;
    int mode;
};

// a search direction (step.jl Point-like Delta); tau / kappa / alpha live in sc[SC_DIR0 + k * DS_STRIDE ...]
struct IpmDir {
    double *x, *xl, *xu, *y, *zl, *zu;
};

struct IpmDev {
    // problem data (ipmdata.jl:14-56): b, c, masked bounds, flags as 0/1 doubles
    const double *b, *c, *lm, *um, *lf, *uf;
    // iterate (point.jl:6-48)
    double *x, *xl, *xu, *y, *zl, *zu;
    // residuals (residuals.jl:6-22)
    double *rp, *rl, *ru, *rd;
    // work vectors of compute_step! (step.jl:24-26, :55-60)
    double *ixl, *ixu, *thl, *thu, *cbar, *hx, *hy, *wl, *wu;
    double* aty_long;   // [nlong] A'y of the long columns (k_ipm_aty_long)
    double* sc;      // [SC_COUNT] device scalars
    double* part;    // [IPM_MAXBLOCKS][IPM_NRED] per-block partials of the fused reductions
};

void ipm_launch_start(const IpmDev& d, int64_t n, int64_t m, double* regP, double* regD, cudaStream_t st);
void ipm_launch_residuals(const IpmDev& d, const DevMat& A, const IpmScalars& P, cudaStream_t st);
void ipm_launch_theta(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, double* theta, double* regP, double* regD,
                      double* xi_d, double* xi_p, cudaStream_t st);
void ipm_launch_bump(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, double* regP, double* regD, cudaStream_t st);
void ipm_launch_h(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, const double* dx, const double* dy, cudaStream_t st);
void ipm_launch_newton_rhs(const IpmDev& d, int64_t n, int64_t m, IpmScalars P, int mode, const IpmDir& D, int slot_dir, double* xi_d,
                           double* xi_p, cudaStream_t st);
void ipm_launch_newton_recover(const IpmDev& d, int64_t n, int64_t m, IpmScalars P, int mode, const IpmDir& D, int slot_dir,
                               const IpmDir& Pv, int slot_prev, const double* dx, const double* dy, cudaStream_t st);
void ipm_launch_targets(const IpmDev& d, int64_t n, const IpmScalars& P, const IpmDir& D, int slot_dir, cudaStream_t st);
void ipm_launch_step(const IpmDev& d, int64_t n, int64_t m, const IpmScalars& P, const IpmDir& D, int slot_dir, cudaStream_t st);

}  // namespace tlp
