// Numeric kernels of the B200 KKT backend (sm_100a): assemble, supernodal signed-Cholesky
// factorisation (K = L S L', S = diag(+-1); S = I for the normal equations), triangular solves.
//
// What they replace in the reference (/root/reference, Tulip.jl v0.9.8):
//   assemble   : kkt.K = A*D*A' + spdiagm(regD)              src/KKT/Cholmod/spd.jl:42-43
//                in-place diagonal update of the K2 matrix   src/KKT/Cholmod/sqd.jl:44-51
//   factor     : cholesky!(F, Symmetric(K)) / ldlt!(F, ...)  spd.jl:46, sqd.jl:53
//   solve      : F \ xi  and the K1 rhs/recovery products    spd.jl:55-66, sqd.jl:61-70
// The dense per-supernode math is what src/KKT/Dense/lapack.jl:85-95,109-110 does on the
// whole matrix.
#include "kernels.cuh"

#include <climits>

namespace tlp {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int32_t pos_in_target(const DevCtx& c, int32_t t, int32_t gi) {
    const int32_t f = c.sn_first[t], l = c.sn_first[t + 1];
    if (gi < l) return gi - f;
    const int64_t rp = c.sn_rowptr[t];
    const int32_t* b = c.sn_rows + rp + (l - f);
    int32_t lo = 0, hi = (int32_t)(c.sn_rowptr[t + 1] - rp) - (l - f);
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(b + mid) < gi) lo = mid + 1; else hi = mid;
    }
    return (l - f) + lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// In-shared signed Cholesky of an nk x nk block stored column-major with leading dimension ldc.
// sg[j] = expected sign of pivot j.  Bad pivots (wrong sign / zero / NaN) are recorded in info
// and replaced by a unit pivot so that the kernel always terminates with finite data.
__device__ void potrf_shared(double* Cs, int ldc, int nk, const double* sg, int32_t* info, int32_t gcol0, int nthr) {
    const int tid = threadIdx.x;
    for (int j = 0; j < nk; ++j) {
        __syncthreads();
        double d = Cs[j * ldc + j];
        const double sj = sg[j];
        if (!(d * sj > 0.0)) {
            if (tid == 0) atomicMin(info, gcol0 + j);
            d = sj;
        }
        const double ljj = sqrt(d * sj);
        const double inv = 1.0 / (sj * ljj);
        __syncthreads();
        for (int i = j + 1 + tid; i < nk; i += nthr) Cs[j * ldc + i] *= inv;
        if (tid == 0) Cs[j * ldc + j] = ljj;
        __syncthreads();
        const int rem = nk - j - 1;
        for (int e = tid; e < rem * rem; e += nthr) {
            const int k = j + 1 + e / rem, i = j + 1 + e % rem;
            if (i >= k) Cs[k * ldc + i] -= Cs[j * ldc + i] * sj * Cs[j * ldc + k];
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// elementwise + assemble
// ------------------------------------------------------------------------------------------
__global__ void k_compute_d(const double* __restrict__ theta, const double* __restrict__ regP, double* __restrict__ d,
                            int64_t n) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < n) d[j] = 1.0 / (theta[j] + regP[j]);   // spd.jl:42  D = inv(Diagonal(theta + regP))
}

// K1: one structural entry of lower(A D A') per thread: sum_j (a_ij a_kj) d_j   (spd.jl:43)
__global__ void k_assemble_k1(double* __restrict__ Lx, const int64_t* __restrict__ w_ptr,
                              const int64_t* __restrict__ w_dest, const int32_t* __restrict__ w_col,
                              const double* __restrict__ w_val, const double* __restrict__ d, int64_t nentries) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nentries) return;
    const int64_t b = w_ptr[e], en = w_ptr[e + 1];
    double s = 0.0;
    for (int64_t p = b; p < en; ++p) s += w_val[p] * __ldg(d + w_col[p]);
    Lx[w_dest[e]] = s;
}

__global__ void k_diag_k1(DevCtx c, const double* __restrict__ regD) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < c.N) c.Lx[c.diagpos[q]] += regD[c.perm[q]];   // + spdiagm(regD)  (spd.jl:43)
}

__global__ void k_scatter_k2(double* __restrict__ Lx, const int64_t* __restrict__ a_dest, const double* __restrict__ val,
                             int64_t nnz) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < nnz) Lx[a_dest[p]] = val[p];
}

__global__ void k_diag_k2(DevCtx c, int64_t n, const double* __restrict__ theta, const double* __restrict__ regP,
                          const double* __restrict__ regD) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= c.N) return;
    const int32_t v = c.perm[q];
    // sqd.jl:44-51: K[j,j] = -theta_j - regP_j ; K[n+i,n+i] = regD_i
    c.Lx[c.diagpos[q]] = (v < n) ? -(theta[v] + regP[v]) : regD[v - n];
}

// ------------------------------------------------------------------------------------------
// small supernodes: whole panel in shared memory, factor + all outgoing updates in one CTA
// ------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 128;

__global__ void __launch_bounds__(SMALL_THREADS) k_small_factor(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    const int32_t s = c.small_list[begin + blockIdx.x];
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int64_t rp = c.sn_rowptr[s];
    const int32_t nr = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    double* Px = c.Lx + c.sn_xptr[s];
    const int tid = threadIdx.x;

    double* Ps = smem_d;                         // nc*nr
    double* sg = Ps + nc * nr;                   // nc
    int32_t* relpos = (int32_t*)(sg + nc);       // nr

    for (int e = tid; e < nc * nr; e += SMALL_THREADS) Ps[e] = Px[e];
    for (int j = tid; j < nc; j += SMALL_THREADS) sg[j] = (double)c.sign[f + j];
    // trapezoid factorisation: potrf of the nc x nc block fused with the scaling of the rows below
    for (int j = 0; j < nc; ++j) {
        __syncthreads();
        double d = Ps[j * nr + j];
        const double sj = sg[j];
        if (!(d * sj > 0.0)) {
            if (tid == 0) atomicMin(c.info, f + j);
            d = sj;
        }
        const double ljj = sqrt(d * sj);
        const double inv = 1.0 / (sj * ljj);
        __syncthreads();
        for (int i = j + 1 + tid; i < nr; i += SMALL_THREADS) Ps[j * nr + i] *= inv;
        if (tid == 0) Ps[j * nr + j] = ljj;
        __syncthreads();
        const int ncr = nc - j - 1, nrr = nr - j - 1;
        for (int e = tid; e < ncr * nrr; e += SMALL_THREADS) {
            const int k = j + 1 + e / nrr, i = j + 1 + e % nrr;
            if (i >= k) Ps[k * nr + i] -= Ps[j * nr + i] * sj * Ps[j * nr + k];
        }
    }
    __syncthreads();
    for (int e = tid; e < nc * nr; e += SMALL_THREADS) Px[e] = Ps[e];

    // outgoing updates, one target segment at a time
    const int64_t g0 = c.seg_ptr[s], g1 = c.seg_ptr[s + 1];
    for (int64_t g = g0; g < g1; ++g) {
        const int32_t kb = c.seg_k0[g];
        const int32_t ke = (g + 1 < g1) ? c.seg_k0[g + 1] : nr;
        const int32_t t = c.seg_tgt[g];
        __syncthreads();
        for (int q = kb + tid; q < nr; q += SMALL_THREADS) relpos[q] = pos_in_target(c, t, rows[q]);
        __syncthreads();
        const int32_t ft = c.sn_first[t];
        const int64_t ldt = c.sn_rowptr[t + 1] - c.sn_rowptr[t];
        double* Tx = c.Lx + c.sn_xptr[t];
        const int nI = nr - kb, nK = ke - kb;
        for (int e = tid; e < nI * nK; e += SMALL_THREADS) {
            const int kk = kb + e / nI, ii = kb + e % nI;
            if (ii < kk) continue;
            double v = 0.0;
            for (int cc = 0; cc < nc; ++cc) v += Ps[cc * nr + ii] * sg[cc] * Ps[cc * nr + kk];
            atomicAdd(Tx + (int64_t)(rows[kk] - ft) * ldt + relpos[ii], -v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// tile update:  C_tgt[pos(I), cols(K)] -= L[I, 0:kdim] * S * L[K, 0:kdim]'   (FP64 DMMA m8n8k4)
// 64x64 output tile per CTA, 4 warps, each warp a 32x32 sub-tile (4x4 mma tiles).
// diag = 1: keep only row >= col.  diag = 2: additionally factor the nk x nk diagonal block.
// ------------------------------------------------------------------------------------------
constexpr int UPD_THREADS = 128;
constexpr int KC = 16;
constexpr int LDT = TILE + 4;   // 68: conflict-free fragment loads (68 mod 16 == 4)
constexpr int LDC = TILE + 1;

__global__ void __launch_bounds__(UPD_THREADS) k_update(DevCtx c, int32_t begin, int atomic) {
    __shared__ double sm[2 * 2 * KC * LDT];      // As[2][KC][LDT], Bs[2][KC][LDT]; reused as Cs[64][65]
    __shared__ int32_t tpos[TILE];
    __shared__ int64_t tcol[TILE];
    __shared__ double sgn[TILE];

    const UpdTask T = c.upd[begin + blockIdx.x];
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    const int32_t f = c.sn_first[s];
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    const double* panel = c.Lx + c.sn_xptr[s] + (int64_t)(pc.c0 - f) * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1, g = lane >> 2, t4 = lane & 3;

    // target addressing
    const int32_t t = T.tgt;
    const int32_t ft = c.sn_first[t];
    const int64_t ldt = c.sn_rowptr[t + 1] - c.sn_rowptr[t];
    if (tid < TILE) {
        int32_t p = 0;
        if (tid < T.ni) p = (t == s) ? (T.i0 + tid) : pos_in_target(c, t, rows[T.i0 + tid]);
        tpos[tid] = p;
    } else {
        const int kk = tid - TILE;
        tcol[kk] = (kk < T.nk) ? (int64_t)(rows[T.k0 + kk] - ft) * ldt : 0;
        sgn[kk] = (kk < T.nk && T.diag == 2) ? (double)c.sign[f + T.k0 + kk] : 1.0;
    }

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    double (*As)[KC][LDT] = reinterpret_cast<double (*)[KC][LDT]>(sm);
    double (*Bs)[KC][LDT] = reinterpret_cast<double (*)[KC][LDT]>(sm + 2 * KC * LDT);

    const int nch = (T.kdim + KC - 1) / KC;
    double ra[8], rb[8];
    auto load_regs = [&](int ch) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = tid + UPD_THREADS * r;
            const int row = e & (TILE - 1), kk = ch * KC + (e >> 6);
            const bool kv = kk < T.kdim;
            const double* col = panel + (int64_t)kk * ld;
            ra[r] = (kv && row < T.ni) ? col[T.i0 + row] : 0.0;
            rb[r] = (kv && row < T.nk) ? col[T.k0 + row] * (double)c.sign[pc.c0 + kk] : 0.0;
        }
    };
    auto store_smem = [&](int st) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = tid + UPD_THREADS * r;
            As[st][e >> 6][e & (TILE - 1)] = ra[r];
            Bs[st][e >> 6][e & (TILE - 1)] = rb[r];
        }
    };
    if (nch > 0) {
        load_regs(0);
        store_smem(0);
    }
    __syncthreads();
    for (int ch = 0; ch < nch; ++ch) {
        const int st = ch & 1;
        if (ch + 1 < nch) load_regs(ch + 1);
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[st][k4 + t4][wr * 32 + mi * 8 + g];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) b[nj] = Bs[st][k4 + t4][wc * 32 + nj * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
        }
        if (ch + 1 < nch) store_smem(st ^ 1);
        __syncthreads();
    }

    double* Tx = c.Lx + c.sn_xptr[t];
    if (T.diag != 2) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ii = wr * 32 + mi * 8 + g, kk = wc * 32 + nj * 8 + t4 * 2 + e;
                    if (ii < T.ni && kk < T.nk && (T.diag == 0 || ii >= kk)) {
                        double* p = Tx + tcol[kk] + tpos[ii];
                        if (atomic) atomicAdd(p, -acc[mi][nj][e]); else *p -= acc[mi][nj][e];
                    }
                }
        return;
    }
    // diag == 2: rows >= nk of the tile are ordinary updates; the nk x nk block is updated in
    // shared memory, factored there and written back (target is the own panel: positions direct)
    double* Cs = sm;
    for (int e = tid; e < T.nk * T.nk; e += UPD_THREADS) {
        const int kk = e / T.nk, ii = e % T.nk;
        Cs[kk * LDC + ii] = (ii >= kk) ? Tx[tcol[kk] + tpos[ii]] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ii = wr * 32 + mi * 8 + g, kk = wc * 32 + nj * 8 + t4 * 2 + e;
                if (ii < T.ni && kk < T.nk && ii >= kk) {
                    if (ii < T.nk) Cs[kk * LDC + ii] -= acc[mi][nj][e];
                    else Tx[tcol[kk] + tpos[ii]] -= acc[mi][nj][e];
                }
            }
    potrf_shared(Cs, LDC, T.nk, sgn, c.info, f + T.k0, UPD_THREADS);
    for (int e = tid; e < T.nk * T.nk; e += UPD_THREADS) {
        const int kk = e / T.nk, ii = e % T.nk;
        if (ii >= kk) Tx[tcol[kk] + tpos[ii]] = Cs[kk * LDC + ii];
    }
}

// ------------------------------------------------------------------------------------------
// trsm: rows below a factored 64-wide diagonal block:  X = A21 * L11^{-T} * S   (one row / thread)
// ------------------------------------------------------------------------------------------
constexpr int TRSM_THREADS = 2 * TILE;

__global__ void __launch_bounds__(TRSM_THREADS) k_trsm(DevCtx c, int32_t begin) {
    __shared__ double Ls[TILE][TILE];    // Ls[j][k] = L11[j,k] * s_k  (k < j)
    __shared__ double invd[TILE];        // 1 / (s_j * L11[j,j])
    const PanelTask T = c.panel[begin + blockIdx.x];
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t kb = (pc.c0 - f) + T.step * TILE;                       // local column of the block
    const int32_t nb = min(TILE, (pc.c1 - pc.c0) - T.step * TILE);
    double* X = c.Lx + c.sn_xptr[s];
    const int tid = threadIdx.x;

    for (int e = tid; e < TILE * TILE; e += TRSM_THREADS) {
        const int j = e / TILE, k = e % TILE;
        double v = 0.0;
        if (j < nb && k < j) v = X[(int64_t)(kb + k) * ld + kb + j] * (double)c.sign[f + kb + k];
        Ls[j][k] = v;
    }
    if (tid < TILE) invd[tid] = (tid < nb) ? 1.0 / ((double)c.sign[f + kb + tid] * X[(int64_t)(kb + tid) * ld + kb + tid]) : 1.0;
    __syncthreads();
    if (tid >= T.nr) return;
    const int32_t r = T.r0 + tid;
    double x[TILE];
#pragma unroll
    for (int j = 0; j < TILE; ++j) x[j] = (j < nb) ? X[(int64_t)(kb + j) * ld + r] : 0.0;
#pragma unroll
    for (int j = 0; j < TILE; ++j) {
        double a = x[j];
#pragma unroll
        for (int k = 0; k < j; ++k) a -= x[k] * Ls[j][k];
        x[j] = a * invd[j];
    }
#pragma unroll
    for (int j = 0; j < TILE; ++j)
        if (j < nb) X[(int64_t)(kb + j) * ld + r] = x[j];
}

// ------------------------------------------------------------------------------------------
// triangular solves  (L u = b ; x = L^{-T} S u), work vector c.wk in permuted order
// ------------------------------------------------------------------------------------------
constexpr int SOLVE_SMALL_WARPS = 4;

__global__ void __launch_bounds__(32 * SOLVE_SMALL_WARPS) k_fwd_small(DevCtx c, int32_t begin, int32_t end) {
    const int lane = threadIdx.x & 31;
    const int32_t idx = begin + blockIdx.x * SOLVE_SMALL_WARPS + (threadIdx.x >> 5);
    if (idx >= end) return;
    const int32_t s = c.small_list[idx];
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int64_t rp = c.sn_rowptr[s];
    const int32_t nr = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    const double* L = c.Lx + c.sn_xptr[s];
    double b = (lane < nc) ? c.wk[f + lane] : 0.0;
    for (int j = 0; j < nc; ++j) {
        const double uj = __shfl_sync(0xffffffffu, b, j) / L[(int64_t)j * nr + j];
        if (lane == j) b = uj;
        else if (lane > j && lane < nc) b -= L[(int64_t)j * nr + lane] * uj;
    }
    if (lane < nc) c.wk[f + lane] = b;
    for (int q0 = nc; q0 < nr; q0 += 32) {
        const int q = q0 + lane;
        double a = 0.0;
        for (int j = 0; j < nc; ++j) {
            const double uj = __shfl_sync(0xffffffffu, b, j);
            if (q < nr) a += L[(int64_t)j * nr + q] * uj;
        }
        if (q < nr) atomicAdd(c.wk + rows[q], -a);
    }
}

__global__ void __launch_bounds__(32 * SOLVE_SMALL_WARPS) k_bwd_small(DevCtx c, int32_t begin, int32_t end) {
    const int lane = threadIdx.x & 31;
    const int32_t idx = begin + blockIdx.x * SOLVE_SMALL_WARPS + (threadIdx.x >> 5);
    if (idx >= end) return;
    const int32_t s = c.small_list[idx];
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int64_t rp = c.sn_rowptr[s];
    const int32_t nr = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    const double* L = c.Lx + c.sn_xptr[s];
    double t = (lane < nc) ? (double)c.sign[f + lane] * c.wk[f + lane] : 0.0;
    for (int q0 = nc; q0 < nr; q0 += 32) {
        const int q = q0 + lane;
        const double xq = (q < nr) ? c.wk[rows[q]] : 0.0;
        for (int j = 0; j < nc; ++j) {
            double v = (q < nr) ? L[(int64_t)j * nr + q] * xq : 0.0;
            v = warp_sum(v);
            if (lane == j) t -= v;
        }
    }
    for (int j = nc - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, t, j) / L[(int64_t)j * nr + j];
        if (lane == j) t = xj;
        else if (lane < j) t -= L[(int64_t)lane * nr + j] * xj;
    }
    if (lane < nc) c.wk[f + lane] = t;
}

// wide pieces: diagonal block (w <= 128) staged in shared memory
constexpr int TRSV_THREADS = 128;

__device__ __forceinline__ void load_diag_block(const DevCtx& c, const Piece& pc, double* Ls, double* invd, int W1) {
    const int32_t s = pc.sn;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t lc0 = pc.c0 - f, w = pc.c1 - pc.c0;
    const double* L = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0;
    for (int e = threadIdx.x; e < w * w; e += blockDim.x) {
        const int k = e / w, i = e % w;
        if (i >= k) Ls[k * W1 + i] = L[(int64_t)k * ld + i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += blockDim.x) invd[j] = 1.0 / Ls[j * W1 + j];
}

__global__ void __launch_bounds__(TRSV_THREADS) k_fwd_trsv(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    const Piece pc = c.pieces[c.level_pieces[begin + blockIdx.x]];
    const int w = pc.c1 - pc.c0, W1 = w + 1;
    double* Ls = smem_d;
    double* invd = Ls + w * W1;
    double* bs = invd + w;
    double* us = bs + w;
    const int tid = threadIdx.x;
    for (int i = tid; i < w; i += TRSV_THREADS) bs[i] = c.wk[pc.c0 + i];
    load_diag_block(c, pc, Ls, invd, W1);
    for (int j = 0; j < w; ++j) {
        __syncthreads();
        const double uj = bs[j] * invd[j];
        if (tid == 0) us[j] = uj;
        for (int i = j + 1 + tid; i < w; i += TRSV_THREADS) bs[i] -= Ls[j * W1 + i] * uj;
    }
    __syncthreads();
    for (int i = tid; i < w; i += TRSV_THREADS) c.wk[pc.c0 + i] = us[i];
}

__global__ void __launch_bounds__(SOLVE_ROWS) k_fwd_gemv(DevCtx c, int32_t begin) {
    __shared__ double us[128];
    const SolveTask T = c.solve[begin + blockIdx.x];
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    const int32_t f = c.sn_first[s];
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int w = pc.c1 - pc.c0;
    const double* L = c.Lx + c.sn_xptr[s] + (int64_t)(pc.c0 - f) * ld;
    const int tid = threadIdx.x;
    if (tid < w) us[tid] = c.wk[pc.c0 + tid];
    __syncthreads();
    if (tid >= T.nr) return;
    const int32_t r = T.r0 + tid;
    double a = 0.0;
#pragma unroll 4
    for (int k = 0; k < w; ++k) a += L[(int64_t)k * ld + r] * us[k];
    atomicAdd(c.wk + c.sn_rows[rp + r], -a);
}

__global__ void __launch_bounds__(SOLVE_ROWS) k_bwd_gemv(DevCtx c, int32_t begin) {
    __shared__ double sacc[128];
    const SolveTask T = c.solve[begin + blockIdx.x];
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    const int32_t f = c.sn_first[s];
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int w = pc.c1 - pc.c0;
    const double* L = c.Lx + c.sn_xptr[s] + (int64_t)(pc.c0 - f) * ld;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 128) sacc[tid] = 0.0;
    __syncthreads();
    const bool valid = tid < T.nr;
    const int32_t r = T.r0 + tid;
    const double xr = valid ? c.wk[c.sn_rows[rp + r]] : 0.0;
    for (int k = 0; k < w; ++k) {
        double v = valid ? L[(int64_t)k * ld + r] * xr : 0.0;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(&sacc[k], v);
    }
    __syncthreads();
    if (tid < w) atomicAdd(c.acc + pc.c0 + tid, sacc[tid]);
}

__global__ void __launch_bounds__(TRSV_THREADS) k_bwd_trsv(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    const Piece pc = c.pieces[c.level_pieces[begin + blockIdx.x]];
    const int w = pc.c1 - pc.c0, W1 = w + 1;
    double* Ls = smem_d;
    double* invd = Ls + w * W1;
    double* ts = invd + w;
    double* xs = ts + w;
    const int tid = threadIdx.x;
    for (int i = tid; i < w; i += TRSV_THREADS) {
        ts[i] = (double)c.sign[pc.c0 + i] * c.wk[pc.c0 + i] - c.acc[pc.c0 + i];
        c.acc[pc.c0 + i] = 0.0;
    }
    load_diag_block(c, pc, Ls, invd, W1);
    for (int j = w - 1; j >= 0; --j) {
        __syncthreads();
        const double xj = ts[j] * invd[j];
        if (tid == 0) xs[j] = xj;
        for (int i = tid; i < j; i += TRSV_THREADS) ts[i] -= Ls[i * W1 + j] * xj;
    }
    __syncthreads();
    for (int i = tid; i < w; i += TRSV_THREADS) c.wk[pc.c0 + i] = xs[i];
}

// ------------------------------------------------------------------------------------------
// right-hand side build / recovery   (spd.jl:55-57, 64-66 ; sqd.jl:61-62, 69-70)
// ------------------------------------------------------------------------------------------
__global__ void k_k1_rhs(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ xi_p,
                         const double* __restrict__ xi_d) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= A.m) return;
    double v = xi_p[i];
    for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
        const int32_t j = A.colidx[p];
        v += A.rval[p] * (d[j] * xi_d[j]);
    }
    c.wk[c.iperm[i]] = v;
}

__global__ void k_k1_recover(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ xi_d,
                             double* __restrict__ dx, double* __restrict__ dy) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < A.n) {
        double v = 0.0;
        for (int64_t p = A.colptr[j]; p < A.colptr[j + 1]; ++p) v += A.val[p] * c.wk[c.iperm[A.rowidx[p]]];
        dx[j] = d[j] * (v - xi_d[j]);
    }
    if (j < A.m) dy[j] = c.wk[c.iperm[j]];
}

__global__ void k_k2_rhs(DevCtx c, DevMat A, const double* __restrict__ xi_p, const double* __restrict__ xi_d) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= A.n + A.m) return;
    c.wk[c.iperm[v]] = (v < A.n) ? xi_d[v] : xi_p[v - A.n];
}

__global__ void k_k2_recover(DevCtx c, DevMat A, double* __restrict__ dx, double* __restrict__ dy) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= A.n + A.m) return;
    const double x = c.wk[c.iperm[v]];
    if (v < A.n) dx[v] = x; else dy[v - A.n] = x;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

void launch_compute_d(const double* theta, const double* regP, double* d, int64_t n, cudaStream_t st) {
    if (n > 0) k_compute_d<<<nblk(n, 256), 256, 0, st>>>(theta, regP, d, n);
}

void launch_assemble_k1(const DevCtx& c, const DevMat& A, const double* d, const double* regD, cudaStream_t st) {
    if (A.nentries > 0)
        k_assemble_k1<<<nblk(A.nentries, 256), 256, 0, st>>>(c.Lx, A.w_ptr, A.w_dest, A.w_col, A.w_val, d, A.nentries);
    if (c.N > 0) k_diag_k1<<<nblk(c.N, 256), 256, 0, st>>>(c, regD);
}

void launch_assemble_k2(const DevCtx& c, const DevMat& A, const double* theta, const double* regP, const double* regD,
                        cudaStream_t st) {
    if (A.nnz > 0) k_scatter_k2<<<nblk(A.nnz, 256), 256, 0, st>>>(c.Lx, A.a_dest, A.val, A.nnz);
    if (c.N > 0) k_diag_k2<<<nblk(c.N, 256), 256, 0, st>>>(c, A.n, theta, regP, regD);
}

size_t small_factor_smem(int32_t max_elems, int32_t max_nrow) {
    return (size_t)max_elems * 8 + 64 * 8 + (size_t)max_nrow * 4 + 16;
}

static size_t trsv_smem(int w) { return ((size_t)w * (w + 1) + 3 * (size_t)w) * 8; }

void kernels_static_init() {
    cudaFuncSetAttribute(k_small_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(k_fwd_trsv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsv_smem(128));
    cudaFuncSetAttribute(k_bwd_trsv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsv_smem(128));
}

void launch_small_factor(const DevCtx& c, int32_t begin, int32_t end, size_t smem, cudaStream_t st) {
    if (end > begin) k_small_factor<<<end - begin, SMALL_THREADS, smem, st>>>(c, begin);
}
void launch_update(const DevCtx& c, int32_t begin, int32_t end, int atomic, cudaStream_t st) {
    if (end > begin) k_update<<<end - begin, UPD_THREADS, 0, st>>>(c, begin, atomic);
}
void launch_trsm(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_trsm<<<end - begin, TRSM_THREADS, 0, st>>>(c, begin);
}
void launch_fwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_fwd_small<<<nblk(end - begin, SOLVE_SMALL_WARPS), 32 * SOLVE_SMALL_WARPS, 0, st>>>(c, begin, end);
}
void launch_bwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_bwd_small<<<nblk(end - begin, SOLVE_SMALL_WARPS), 32 * SOLVE_SMALL_WARPS, 0, st>>>(c, begin, end);
}
void launch_fwd_trsv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_fwd_trsv<<<end - begin, TRSV_THREADS, trsv_smem(128), st>>>(c, begin);
}
void launch_bwd_trsv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_bwd_trsv<<<end - begin, TRSV_THREADS, trsv_smem(128), st>>>(c, begin);
}
void launch_fwd_gemv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_fwd_gemv<<<end - begin, SOLVE_ROWS, 0, st>>>(c, begin);
}
void launch_bwd_gemv(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_bwd_gemv<<<end - begin, SOLVE_ROWS, 0, st>>>(c, begin);
}
void launch_k1_rhs(const DevCtx& c, const DevMat& A, const double* d, const double* xi_p, const double* xi_d,
                   cudaStream_t st) {
    if (A.m > 0) k_k1_rhs<<<nblk(A.m, 128), 128, 0, st>>>(c, A, d, xi_p, xi_d);
}
void launch_k1_recover(const DevCtx& c, const DevMat& A, const double* d, const double* xi_d, double* dx, double* dy,
                       cudaStream_t st) {
    const int64_t mx = A.n > A.m ? A.n : A.m;
    if (mx > 0) k_k1_recover<<<nblk(mx, 128), 128, 0, st>>>(c, A, d, xi_d, dx, dy);
}
void launch_k2_rhs(const DevCtx& c, const DevMat& A, const double* xi_p, const double* xi_d, cudaStream_t st) {
    if (A.n + A.m > 0) k_k2_rhs<<<nblk(A.n + A.m, 256), 256, 0, st>>>(c, A, xi_p, xi_d);
}
void launch_k2_recover(const DevCtx& c, const DevMat& A, double* dx, double* dy, cudaStream_t st) {
    if (A.n + A.m > 0) k_k2_recover<<<nblk(A.n + A.m, 256), 256, 0, st>>>(c, A, dx, dy);
}

}  // namespace tlp
