// Dense-column handling for the normal equations (BASELINE config 5; a capability the reference lacks:
// with K1 its A*D*A' simply becomes dense, /root/reference/src/KKT/Cholmod/spd.jl:43).
//
//   A = [A_s  A_d],  K = A_s D_s A_s' + Rd + A_d D_d A_d' = K_s + A_d D_d A_d',   K_s = L L'
// K_s keeps the sparse factorisation; the nd dense columns enter through the Schur complement of the bordered system, in
// FACTORISED form (block elimination of [K_s A_d; A_d' -D_d^{-1}] with the sparse part first):
//   W = L^{-1} A_d                 (nd forward sweeps per update!)
//   C = D_d^{-1} + W'W             (nd x nd, a Gram matrix + positive diagonal: SPD by construction), C = Lc Lc'
//   K^{-1} b = L^{-T} ( z - W C^{-1} W'z ),   z = L^{-1} b
// Round 1 used the explicit Sherman-Morrison-Woodbury formula y0 - V C^{-1} A_d'y0 with V = K_s^{-1} A_d, y0 = K_s^{-1} b.
// Near IPM convergence the dense columns are basic and K_s alone is singular up to the regularisation (cond ~ 1e16): both
// terms of that formula are amplified by 1/lambda_min(K_s) and cancel, and the refinement built on it DIVERGES -- the
// config-5 IPM stalled at pfeas ~ 5 from iteration 18 on (reproduced with a NumPy emulation on the CPU port's factor;
// 17 iterations to Trm_Optimal with the factorised form, the same count as the CPU port's K2).  In the factorised form the
// correction is subtracted half-way, before L^{-T} amplifies anything.
//
// The forward-sweep kernels leave the blocks of a dense-solve ("big") supernode scaled by its diagonal blocks
// (kernels_dense_solve.cu: w = G u, G = blockdiag(L_kk) there, identity elsewhere).  With Wt = G W (what the sweep
// returns) and Wh = (G G')^{-1} Wt:  W'W = Wh'Wt,  W'z = Wh'(G z),  G (z - W t) = G z - Wt t -- so everything is done on
// the vectors exactly as the sweeps produce and consume them.
#include "kernels.cuh"

namespace tlp {

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

// wk = A_d(:, j) in permuted row order (wk must be zero on entry)
__global__ void k_dc_scatter(DevCtx c, DenseCols dc, int j) {
    const int64_t p = dc.colptr[j] + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < dc.colptr[j + 1]) c.wk[dc.prow[p]] = dc.val[p];
}

// C[i][j] = Wh_i' Wt_j  (+ theta+regP of dense column i on the diagonal = D_d^{-1})
__global__ void k_dc_gram(DevCtx c, DenseCols dc, const double* __restrict__ theta, const double* __restrict__ regP) {
    __shared__ double red[32];
    const int i = blockIdx.x, j = blockIdx.y;
    const double* Wi = dc.Wh + (int64_t)i * c.N;
    const double* Wj = dc.Wt + (int64_t)j * c.N;
    double a = 0.0;
    for (int32_t q = threadIdx.x; q < c.N; q += blockDim.x) a += Wi[q] * Wj[q];
    a = block_sum(a, red);
    if (threadIdx.x == 0) {
        if (i == j) a += theta[dc.col_id[i]] + regP[dc.col_id[i]];
        dc.C[i * dc.nd + j] = a;
    }
}

// out = (G G')^{-1} in on the diagonal blocks of the dense-solve supernodes (G_kk = L_kk): two triangular products with the
// explicit inverse X = L_kk^{-1} (Dinv: column-major lower, DinvT: its transpose); one CTA per block.  `out` already holds a
// copy of `in` for all other entries.
__global__ void __launch_bounds__(SBLK) k_dc_ginv(DevCtx c, DenseCols dc, const double* __restrict__ in, double* __restrict__ out) {
    __shared__ double xs[SBLK], ys[SBLK];
    const int32_t d = dc.gblk[3 * blockIdx.x], col0 = dc.gblk[3 * blockIdx.x + 1], w = dc.gblk[3 * blockIdx.x + 2];
    const double* X = c.Dinv + (int64_t)d * SBLK * SBLK;      // X[r, cc] at cc * SBLK + r
    const double* XT = c.DinvT + (int64_t)d * SBLK * SBLK;    // X[r, cc] at r * SBLK + cc
    const int t = threadIdx.x;
    xs[t] = t < w ? in[col0 + t] : 0.0;
    __syncthreads();
    double y = 0.0;
    if (t < w)
        for (int cc = 0; cc <= t; ++cc) y += X[cc * SBLK + t] * xs[cc];        // y = L^{-1} x
    ys[t] = y;
    __syncthreads();
    if (t < w) {
        double z = 0.0;
        for (int r = t; r < w; ++r) z += XT[r * SBLK + t] * ys[r];             // z = L^{-T} y
        out[col0 + t] = z;
    }
}

// in-place Cholesky of the nd x nd matrix C (one block; nd <= 64)
__global__ void k_dc_chol(DevCtx c, DenseCols dc) {
    __shared__ double Cs[64 * 65];
    const int nd = dc.nd, tid = threadIdx.x;
    for (int e = tid; e < nd * nd; e += blockDim.x) Cs[(e / nd) * 65 + (e % nd)] = 0.5 * (dc.C[e] + dc.C[(e % nd) * nd + e / nd]);
    for (int j = 0; j < nd; ++j) {
        __syncthreads();
        double d = Cs[j * 65 + j];
        if (!(d > 0.0)) {
            if (tid == 0) atomicMin(c.info, c.N - 1);
            d = 1.0;
        }
        const double l = sqrt(d);
        __syncthreads();
        for (int i = j + tid; i < nd; i += blockDim.x) Cs[i * 65 + j] = (i == j) ? l : Cs[i * 65 + j] / l;
        __syncthreads();
        for (int e = tid; e < (nd - j - 1) * (nd - j - 1); e += blockDim.x) {
            const int i = j + 1 + e / (nd - j - 1), k = j + 1 + e % (nd - j - 1);
            if (i >= k) Cs[i * 65 + k] -= Cs[i * 65 + j] * Cs[k * 65 + j];
        }
    }
    __syncthreads();
    for (int e = tid; e < nd * nd; e += blockDim.x) {
        const int i = e / nd, k = e % nd;
        dc.C[e] = (i >= k) ? Cs[i * 65 + k] : 0.0;     // row-major lower factor Lc
    }
}

// g[i] = Wh_i' wk   (= W_i' z for the forward-swept vector in wk)
__global__ void k_dc_dots(DevCtx c, DenseCols dc) {
    __shared__ double red[32];
    const int i = blockIdx.x;
    const double* Wi = dc.Wh + (int64_t)i * c.N;
    double a = 0.0;
    for (int32_t q = threadIdx.x; q < c.N; q += blockDim.x) a += Wi[q] * c.wk[q];
    a = block_sum(a, red);
    if (threadIdx.x == 0) dc.g[i] = a;
}

// t = C^{-1} g (every block redundantly, nd <= 64), then wk -= sum_i Wt_i t_i
__global__ void k_dc_apply(DevCtx c, DenseCols dc) {
    __shared__ double t[64];
    const int nd = dc.nd;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nd; ++i) {                      // Lc y = g
            double a = dc.g[i];
            for (int k = 0; k < i; ++k) a -= dc.C[i * nd + k] * t[k];
            t[i] = a / dc.C[i * nd + i];
        }
        for (int i = nd - 1; i >= 0; --i) {                 // Lc' t = y
            double a = t[i];
            for (int k = i + 1; k < nd; ++k) a -= dc.C[k * nd + i] * t[k];
            t[i] = a / dc.C[i * nd + i];
        }
    }
    __syncthreads();
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= c.N) return;
    double a = 0.0;
    for (int i = 0; i < nd; ++i) a += dc.Wt[(int64_t)i * c.N + q] * t[i];
    c.wk[q] -= a;
}

// ---- iterative refinement on the full normal equations (the Woodbury correction alone loses accuracy when K_s is
// ill-conditioned, which it is near IPM convergence):  r = xi - (A D A' + Rd) y, all in permuted row order
__global__ void k_dc_at_y(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ y, double* __restrict__ tn) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    const int64_t b = A.colptr[j], e = A.colptr[j + 1];
    if (e - b > LONG_COL) return;            // one CTA per long column: k_dc_at_y_long
    double v = 0.0;
    for (int64_t p = b; p < e; ++p) v += A.val[p] * y[c.iperm[A.rowidx[p]]];
    tn[j] = d[j] * v;
}
__global__ void __launch_bounds__(256) k_dc_at_y_long(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ y,
                                                     double* __restrict__ tn) {
    __shared__ double red[32];
    const int64_t j = A.long_cols[blockIdx.x];
    double v = 0.0;
    for (int64_t p = A.colptr[j] + threadIdx.x; p < A.colptr[j + 1]; p += blockDim.x) v += A.val[p] * y[c.iperm[A.rowidx[p]]];
    v = block_sum(v, red);
    if (threadIdx.x == 0) tn[j] = d[j] * v;
}
__global__ void k_dc_residual(DevCtx c, DevMat A, const double* __restrict__ regD, const double* __restrict__ xi,
                              const double* __restrict__ y, const double* __restrict__ tn) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= A.m) return;
    const int32_t q = c.iperm[i];
    double v = xi[q] - regD[i] * y[q];
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) v -= A.rval[p] * tn[A.colidx[p]];
    c.wk[q] = v;
}
__global__ void k_dc_axpy(DevCtx c, const double* __restrict__ y) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < c.N) c.wk[q] += y[q];
}
// K2: r = [xi_d; xi_p] - [-(theta + Rp) A'; A Rd] [dx; dy]  (systems.jl:8-32 signs), everything in permuted order:
// xi / y are the saved right-hand side and the current solution as laid out in wk (index iperm[v], v < n: x-block)
__global__ void k_k2_residual(DevCtx c, DevMat A, const double* __restrict__ theta, const double* __restrict__ regP,
                              const double* __restrict__ regD, const double* __restrict__ xi, const double* __restrict__ y) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= A.n + A.m) return;
    const int32_t q = c.iperm[v];
    double r = xi[q];
    if (v < A.n) {
        r += (theta[v] + regP[v]) * y[q];
        for (int64_t p = A.colptr[v]; p < A.colptr[v + 1]; ++p) r -= A.val[p] * y[c.iperm[A.n + A.rowidx[p]]];
    } else {
        const int64_t i = v - A.n;
        r -= regD[i] * y[q];
        for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) r -= A.rval[p] * y[c.iperm[A.colidx[p]]];
    }
    c.wk[q] = r;
}
void launch_k2_residual(const DevCtx& c, const DevMat& A, const double* theta, const double* regP, const double* regD, const double* xi,
                        const double* y, cudaStream_t st) {
    if (A.n + A.m > 0) k_k2_residual<<<(unsigned)((A.n + A.m + 127) / 128), 128, 0, st>>>(c, A, theta, regP, regD, xi, y);
}

void launch_dc_residual(const DevCtx& c, const DevMat& A, const double* d, const double* regD, const double* xi, const double* y,
                        double* tn, cudaStream_t st) {
    k_dc_at_y<<<(unsigned)((A.n + 127) / 128), 128, 0, st>>>(c, A, d, y, tn);
    if (A.nlong > 0) k_dc_at_y_long<<<A.nlong, 256, 0, st>>>(c, A, d, y, tn);
    k_dc_residual<<<(unsigned)((A.m + 127) / 128), 128, 0, st>>>(c, A, regD, xi, y, tn);
}
void launch_dc_axpy(const DevCtx& c, const double* y, cudaStream_t st) {
    k_dc_axpy<<<(c.N + 255) / 256, 256, 0, st>>>(c, y);
}

void launch_dc_scatter(const DevCtx& c, const DenseCols& dc, int j, int64_t colnnz, cudaStream_t st) {
    if (colnnz > 0) k_dc_scatter<<<(unsigned)((colnnz + 255) / 256), 256, 0, st>>>(c, dc, j);
}
void launch_dc_gram_chol(const DevCtx& c, const DenseCols& dc, const double* theta, const double* regP, cudaStream_t st) {
    k_dc_gram<<<dim3(dc.nd, dc.nd), 256, 0, st>>>(c, dc, theta, regP);
    k_dc_chol<<<1, 256, 0, st>>>(c, dc);
}
void launch_dc_ginv(const DevCtx& c, const DenseCols& dc, const double* in, double* out, cudaStream_t st) {
    cudaMemcpyAsync(out, in, (size_t)c.N * 8, cudaMemcpyDeviceToDevice, st);
    if (dc.ngblk > 0) k_dc_ginv<<<dc.ngblk, SBLK, 0, st>>>(c, dc, in, out);
}
void launch_dc_apply(const DevCtx& c, const DenseCols& dc, cudaStream_t st) {
    k_dc_dots<<<dc.nd, 256, 0, st>>>(c, dc);
    k_dc_apply<<<(c.N + 255) / 256, 256, 0, st>>>(c, dc);
}

}  // namespace tlp
