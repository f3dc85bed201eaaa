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
#include <cstdlib>

namespace tlp {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int32_t ld_acquire(const int32_t* p) {
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int32_t v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

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
    if (c.skip && c.skip[s]) return;
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
// explicit inverses of the 128x128 diagonal blocks (used by the dense solve kernels):
// thread c computes column c of X = L^{-1} by forward substitution.
// ------------------------------------------------------------------------------------------
constexpr int INV_THREADS = SBLK;

__global__ void __launch_bounds__(INV_THREADS) k_invert_diag(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Cs = smem_d;                 // [SBLK][SBLK+1] column-major: L on entry, X = L^{-1} on exit
    const int LDI = SBLK + 1;
    const int32_t b = c.inv_order[begin + blockIdx.x];
    const int32_t s = c.dblk_sn[b], bi = c.dblk_idx[b];
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t lc0 = bi * SBLK, nb = min(SBLK, nc - lc0);
    const double* D = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0;
    const int tid = threadIdx.x;
    for (int k = 0; k < nb; ++k)
        if (tid < nb) Cs[k * LDI + tid] = (tid >= k) ? D[(int64_t)k * ld + tid] : 0.0;
    // in-place inverse of a lower-triangular matrix, last column first:
    // X[j,j] = 1/L[j,j];  X[i,j] = -X[j,j] * sum_{k=j+1..i} X[i,k] L[k,j]   (i > j)
    for (int j = nb - 1; j >= 0; --j) {
        __syncthreads();
        const double xjj = 1.0 / Cs[j * LDI + j];
        double v = 0.0;
        if (tid > j && tid < nb) {
            double a = 0.0;
            for (int k = j + 1; k <= tid; ++k) a += Cs[k * LDI + tid] * Cs[j * LDI + k];
            v = -a * xjj;
        }
        __syncthreads();
        if (tid > j && tid < nb) Cs[j * LDI + tid] = v;
        if (tid == j) Cs[j * LDI + j] = xjj;
    }
    __syncthreads();
    double* out = c.Dinv + (int64_t)b * SBLK * SBLK;
    double* outT = c.DinvT + (int64_t)b * SBLK * SBLK;
    for (int e = tid; e < SBLK * SBLK; e += INV_THREADS) {
        const int cc = e / SBLK, r = e % SBLK;
        out[cc * SBLK + r] = (cc < nb && r < nb && r >= cc) ? Cs[cc * LDI + r] : 0.0;     // Dinv  (column cc, row r)
    }
    for (int e = tid; e < SBLK * SBLK; e += INV_THREADS) {
        const int r = e / SBLK, cc = e % SBLK;
        outT[r * SBLK + cc] = (cc < nb && r < nb && r >= cc) ? Cs[cc * LDI + r] : 0.0;    // DinvT (column r, row cc) = X[r, cc]
    }
    // transposed copy of the sub-diagonal tile L[(bi+1) block rows, bi block cols] for the backward sweep:
    // LsubT[b][rr * 128 + cc] = L[lc0 + 128 + rr, lc0 + cc]
    __syncthreads();                                         // Cs is reused as the transposition buffer
    double* sub = c.LsubT + (int64_t)b * SBLK * SBLK;
    const int32_t nbn = min(SBLK, nc - (lc0 + SBLK));       // rows of the next diagonal block (<= 0: none)
    const double* Lsub = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0 + SBLK;
    for (int cc = 0; cc < SBLK; ++cc) {
        const int rr = tid;                                  // coalesced read along the rows
        Cs[cc * LDI + rr] = (rr < nbn && cc < nb) ? Lsub[(int64_t)cc * ld + rr] : 0.0;
    }
    __syncthreads();
    for (int e = tid; e < SBLK * SBLK; e += INV_THREADS) {
        const int rr = e / SBLK, cc = e % SBLK;
        sub[rr * SBLK + cc] = Cs[cc * LDI + rr];
    }
}

// ------------------------------------------------------------------------------------------
// k_invert_diag2 (round 2): the same three outputs (Dinv, DinvT, LsubT) with the inverse built by block doubling on the FP64
// tensor pipe.  k_invert_diag lets thread i run a dependent dot product of length i - j for every column j (8 128 chained
// FMAs for the last row, two barriers per column): 270 us per block, 1.4 ms of config 4's 8.3 ms update!.  Here:
//   1. the eight 16x16 diagonal blocks by substitution, one thread per column (128 threads, reciprocal diagonal precomputed);
//   2. for h = 16, 32, 64: L = [A 0; B C] with A^-1, C^-1 known  ->  T = B A^-1, X21 = -C^-1 T, both as DMMA 8x8x4 tile
//      products out of shared memory (operands column-major, leading dimension 132: conflict-free fragments), zero tiles of
//      the triangular factors skipped; 3 x 2 barriers instead of 256;
//   3. Dinv with 16-byte stores, DinvT through a 128 x 32 staging buffer of odd stride (no bank conflicts on the transposed
//      read), LsubT as before.
// ------------------------------------------------------------------------------------------
constexpr int INV2_THREADS = 512;
constexpr int LDI2 = SBLK + 4;      // 132
constexpr int LDT2 = 68;            // T = B A^-1, at most 64 x 64
constexpr int LDS2 = 33;            // transposition staging [128][33]

// tiles (mi, nj0 .. nj0 + TPW - 1) of pair p at level h: GEMM 1 (T = B A^-1) or GEMM 2 (X21 = -C^-1 T)
template <int TPW, bool SECOND>
__device__ __forceinline__ void inv2_tiles(double* Cs, double* Tb, int h, int p, int mi, int nj0, int g, int t4) {
    const int a0 = 2 * p * h, c0 = a0 + h;      // first row / column of A and of C
    double acc[TPW][2];
#pragma unroll
    for (int x = 0; x < TPW; ++x) acc[x][0] = acc[x][1] = 0.0;
    // GEMM 1: A^-1 is lower triangular: rows k < 8 nj of its column tile nj are zero.  GEMM 2: C^-1 lower: columns k >= 8 (mi + 1)
    // of its row tile mi are zero.
    const int kbeg = SECOND ? 0 : 8 * nj0, kend = SECOND ? 8 * (mi + 1) : h;
    for (int k0 = kbeg; k0 < kend; k0 += 4) {
        const double a = SECOND ? Cs[(c0 + k0 + t4) * LDI2 + c0 + 8 * mi + g]      // C^-1[row, k]
                                : Cs[(a0 + k0 + t4) * LDI2 + c0 + 8 * mi + g];     // B[row, k]
#pragma unroll
        for (int x = 0; x < TPW; ++x) {
            const int nj = nj0 + x;
            const double b = SECOND ? Tb[(p * h + 8 * nj + g) * LDT2 + k0 + t4]          // T[k, n]
                                    : Cs[(a0 + 8 * nj + g) * LDI2 + a0 + k0 + t4];       // A^-1[k, n]
            dmma884(acc[x][0], acc[x][1], a, b);
        }
    }
#pragma unroll
    for (int x = 0; x < TPW; ++x) {
        const int nj = nj0 + x;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (SECOND) Cs[(a0 + 8 * nj + 2 * t4 + e) * LDI2 + c0 + 8 * mi + g] = -acc[x][e];      // over B, which GEMM 1 has consumed
            else Tb[(p * h + 8 * nj + 2 * t4 + e) * LDT2 + 8 * mi + g] = acc[x][e];
        }
    }
}

template <int H>
__device__ __forceinline__ void inv2_level(double* Cs, double* Tb, int warp, int g, int t4) {
    constexpr int TPW = H / 16;               // tiles per warp: 16 warps share 64 / H pairs x (H / 8)^2 tiles
    constexpr int WPP = H / 4;                // warps per pair
    const int p = warp / WPP, widx = warp % WPP;
    const int mi = widx >> 1, nj0 = (widx & 1) * TPW;
    inv2_tiles<TPW, false>(Cs, Tb, H, p, mi, nj0, g, t4);
    __syncthreads();
    inv2_tiles<TPW, true>(Cs, Tb, H, p, mi, nj0, g, t4);
    __syncthreads();
}

__global__ void __launch_bounds__(INV2_THREADS, 1) k_invert_diag2(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Cs = smem_d;                      // [SBLK][LDI2] column-major: L on entry, X = L^{-1} on exit (upper part zero)
    double* Tb = Cs + SBLK * LDI2;            // [64][LDT2] / [128][16] / [128][LDS2] scratch
    double* rdiag = Tb + 64 * LDT2;           // [SBLK]
    const int32_t b = c.inv_order[begin + blockIdx.x];
    const int32_t s = c.dblk_sn[b], bi = c.dblk_idx[b];
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t lc0 = bi * SBLK, nb = min(SBLK, nc - lc0);
    const double* D = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int il = tid & (SBLK - 1), q4 = tid >> 7;
    // lower triangle in; zero above it; identity on the padding rows / columns (nb < 128)
    for (int kb = 0; kb < SBLK; kb += 64) {
        double v[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = kb + q4 + 4 * x;
            v[x] = (k < nb && il < nb && il >= k) ? D[(int64_t)k * ld + il] : ((k == il && k >= nb) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) Cs[(kb + q4 + 4 * x) * LDI2 + il] = v[x];
    }
    __syncthreads();
    if (tid < SBLK) rdiag[tid] = 1.0 / Cs[tid * LDI2 + tid];
    __syncthreads();
    // 1. 16x16 diagonal blocks: thread t = column t of the block inverse, x_i = (delta_ic - sum_{k<i} L_ik x_k) / L_ii
    if (tid < SBLK) {
        const int d0 = tid & ~15, cc = tid & 15;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double a = (i == cc) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) a = fma(-Cs[(d0 + k) * LDI2 + d0 + i], x[k], a);
            x[i] = a * rdiag[d0 + i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) Tb[tid * 16 + i] = x[i];
    }
    __syncthreads();
    if (tid < SBLK) {
        const int d0 = tid & ~15;
#pragma unroll
        for (int i = 0; i < 16; ++i) Cs[tid * LDI2 + d0 + i] = Tb[tid * 16 + i];
    }
    __syncthreads();
    // 2. block doubling
    inv2_level<16>(Cs, Tb, warp, g, t4);
    inv2_level<32>(Cs, Tb, warp, g, t4);
    inv2_level<64>(Cs, Tb, warp, g, t4);
    // 3. outputs.  Dinv (column cc, row r), 16-byte stores
    double* out = c.Dinv + (int64_t)b * SBLK * SBLK;
    double* outT = c.DinvT + (int64_t)b * SBLK * SBLK;
    for (int e = tid; e < SBLK * SBLK / 2; e += INV2_THREADS) {
        const int cc = e >> 6, r = (e & 63) * 2;
        const double2 v = *reinterpret_cast<const double2*>(Cs + cc * LDI2 + r);
        double2 o;
        o.x = (cc < nb && r < nb && r >= cc) ? v.x : 0.0;
        o.y = (cc < nb && r + 1 < nb && r + 1 >= cc) ? v.y : 0.0;
        *reinterpret_cast<double2*>(out + cc * SBLK + r) = o;
    }
    // DinvT (column r, row cc) = X[r, cc]: 32 columns cc at a time through Tb[r][LDS2]
    for (int cb = 0; cb < SBLK; cb += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * SBLK; e += INV2_THREADS) {
            const int cc = cb + (e >> 7), r = e & (SBLK - 1);
            Tb[r * LDS2 + (cc - cb)] = (cc < nb && r < nb && r >= cc) ? Cs[cc * LDI2 + r] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < 32 * SBLK; e += INV2_THREADS) {
            const int r = e >> 5, j = e & 31;
            outT[r * SBLK + cb + j] = Tb[r * LDS2 + j];
        }
    }
    // transposed copy of the sub-diagonal tile L[(bi+1) block rows, bi block cols] for the backward sweep:
    // LsubT[b][rr * 128 + cc] = L[lc0 + 128 + rr, lc0 + cc]; Cs is reused as [128][129]
    __syncthreads();
    constexpr int LDO = SBLK + 1;
    double* sub = c.LsubT + (int64_t)b * SBLK * SBLK;
    const int32_t nbn = min(SBLK, nc - (lc0 + SBLK));       // rows of the next diagonal block (<= 0: none)
    const double* Lsub = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0 + SBLK;
    for (int kb = 0; kb < SBLK; kb += 64) {
        double v[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int cc = kb + q4 + 4 * x;
            v[x] = (il < nbn && cc < nb) ? Lsub[(int64_t)cc * ld + il] : 0.0;
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) Cs[(kb + q4 + 4 * x) * LDO + il] = v[x];
    }
    __syncthreads();
    for (int e = tid; e < SBLK * SBLK; e += INV2_THREADS) {
        const int rr = e / SBLK, cc = e % SBLK;
        sub[rr * SBLK + cc] = Cs[cc * LDO + rr];
    }
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
    if (c.skip && c.skip[s]) return;
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
    if (c.skip && c.skip[s]) return;
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

// ---- dense block solve of the non-small supernodes ----------------------------------------------
// Persistent CTAs (<= one per SM, all co-resident) walk a wavefront-ordered item list; a diagonal
// block publishes its solution through a release/acquire flag, consumers spin on it.  Each item
// streams its 128x128 tiles of L exactly once (the HBM-roofline traffic of a triangular solve).
constexpr int SL_THREADS = 512;
constexpr int SL_CG = SL_THREADS / SBLK;   // 4 column groups of 32

// The spin is bounded: the grid is sized for co-residency, but anything else holding SMs (a second handle, a
// caller's stream, NCCL) could keep a producer CTA off the device; a wait that runs out reports through info[2]
// (-> TLPB200_INTERNAL from the solve) instead of hanging the GPU.
__device__ __forceinline__ void wait_flag(const int32_t* flag, int32_t* info) {
    if (threadIdx.x == 0) {
        int spins = 0;
        while (ld_acquire(flag) == 0) {
            if (++spins > (1 << 22)) { atomicExch(info + 2, 1); break; }
        }
    }
    __syncthreads();
}

// bounded wait until *cnt >= need (merged-level sweeps: "all items of my in-launch children / parent have finished")
__device__ __forceinline__ void wait_count(const int32_t* cnt, int32_t need, int32_t* info) {
    if (threadIdx.x == 0) {
        int spins = 0;
        while (ld_acquire(cnt) < need) {
            if (++spins > (1 << 22)) { atomicExch(info + 2, 1); break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SL_THREADS, 1) k_fwd_large(DevCtx c, int32_t begin, int32_t end, int merged) {
    extern __shared__ double smem_d[];
    double* Ds = smem_d;                   // [SBLK*SBLK] explicit inverse of the item's diagonal block
    double* xs = Ds + SBLK * SBLK;         // [SBLK]
    double* red = xs + SBLK;               // [SL_CG][SBLK]
    const int tid = threadIdx.x, r = tid & (SBLK - 1), cg = tid >> 7;
    int32_t* fflag = c.flags;
    for (int32_t it = begin + blockIdx.x; it < end; it += gridDim.x) {
        const SolveItem I = c.fwd_items[it];
        const int32_t s = I.sn;
        if (c.skip && c.skip[s]) continue;
        const int32_t f = c.sn_first[s];
        const int32_t nc = c.sn_first[s + 1] - f;
        const int64_t rp = c.sn_rowptr[s];
        const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
        const double* panel = c.Lx + c.sn_xptr[s];
        const int32_t db = c.sn_dblk[s];
        const int32_t ncb = (nc + SBLK - 1) / SBLK;
        const int32_t njt = (I.kind == 0) ? I.blk : ncb;
        const bool rvalid = r < I.nr;
        const int32_t prow = I.r0 + (rvalid ? r : 0);
        if (I.kind == 0) {   // stage the inverse of the diagonal block (needed last)
            const double* src = c.Dinv + (int64_t)(db + I.blk) * SBLK * SBLK;
            for (int e = tid; e < SBLK * SBLK / 2; e += SL_THREADS) cp_async16(Ds + 2 * e, src + 2 * e);
            cp_async_commit();
        }
        double t[32];
        double acc = 0.0;
        auto load_tile = [&](int j) {
            const int32_t nbj = min(SBLK, nc - j * SBLK);
            const double* col = panel + (int64_t)(j * SBLK + cg * 32) * ld + prow;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) t[cc] = (rvalid && cg * 32 + cc < nbj) ? col[(int64_t)cc * ld] : 0.0;
        };
        if (njt > 0) load_tile(0);
        for (int j = 0; j < njt; ++j) {
            wait_flag(fflag + db + j, c.info);
            if (tid < SBLK) xs[tid] = (j * SBLK + tid < nc) ? __ldcg(c.wk + f + j * SBLK + tid) : 0.0;
            __syncthreads();
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) acc -= t[cc] * xs[cg * 32 + cc];
            if (j + 1 < njt) load_tile(j + 1);
            __syncthreads();      // xs is rewritten in the next round
        }
        red[cg * SBLK + r] = acc;
        __syncthreads();
        // merged levels: the right-hand side of this supernode's columns is final once every item of its children inside
        // this launch has finished (children outside the launch finished before it started).  It is only consumed here, after
        // the tiles of the earlier blocks of the same supernode: for all but a supernode's first block the wait is off the chain.
        if (merged && I.kind == 0 && c.fwd_need[s] > 0) wait_count(c.dep_cnt + s, c.fwd_need[s], c.info);
        const double bown = (I.kind == 0 && tid < SBLK && rvalid) ? __ldcg(c.wk + f + I.r0 + tid) : 0.0;
        if (I.kind == 1) {
            if (tid < SBLK && rvalid) {
                const double v = red[r] + red[SBLK + r] + red[2 * SBLK + r] + red[3 * SBLK + r];
                atomicAdd(c.wk + c.sn_rows[rp + I.r0 + r], v);
            }
            if (merged && c.fwd_parent[s] >= 0) {      // publish: this item's contributions are in place
                __threadfence();
                __syncthreads();
                if (tid == 0) atomicAdd(c.dep_cnt + c.fwd_parent[s], 1);
            }
            __syncthreads();
            continue;
        }
        if (tid < SBLK)
            xs[tid] = rvalid ? (bown + red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid]) : 0.0;
        cp_async_wait_all();
        __syncthreads();
        double a2 = 0.0;
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) a2 += Ds[(cg * 32 + cc) * SBLK + r] * xs[cg * 32 + cc];
        __syncthreads();
        red[cg * SBLK + r] = a2;
        __syncthreads();
        if (tid < SBLK && rvalid)
            __stcg(c.wk + f + I.r0 + tid, red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            st_release(fflag + db + I.blk, 1);
            if (merged && c.fwd_parent[s] >= 0) atomicAdd(c.dep_cnt + c.fwd_parent[s], 1);
        }
    }
}

__device__ __forceinline__ void bwd_below_item(const DevCtx& c, int32_t s, int32_t i, int32_t br0, int32_t bnr, double* red) {
    const int tid = threadIdx.x, r = tid & (SBLK - 1), cg = tid >> 7, lane = tid & 31, wq = (tid >> 5) & 3;
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t nbi = min(SBLK, nc - i * SBLK);
    const double* colbase = c.Lx + c.sn_xptr[s] + (int64_t)(i * SBLK + cg * 32) * ld;
    const int32_t ncv = min(32, max(0, nbi - cg * 32));
    double p[32];
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) p[cc] = 0.0;
    const int32_t rend = br0 + bnr;
#pragma unroll 1
    for (int32_t r0 = br0; r0 < rend; r0 += SBLK) {
        const int32_t rr = r0 + r;
        if (rr < rend) {
            const double xr = __ldcg(c.wk + c.sn_rows[rp + rr]);
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8) {   // 8 loads in flight per thread (keeps the kernel out of local memory)
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = (c0 + q < ncv) ? __ldcs(colbase + (int64_t)(c0 + q) * ld + rr) : 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) p[c0 + q] += v[q] * xr;
                asm volatile("" ::: "memory");
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const double send = up ? p[k] : p[k + off];
            const double keep = up ? p[k + off] : p[k];
            p[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    red[wq * SBLK + cg * 32 + lane] = p[0];
    __syncthreads();
    if (tid < nbi) atomicAdd(c.bacc + f + i * SBLK + tid, red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid]);
}

__global__ void __launch_bounds__(SL_THREADS, 1) k_bwd_large(DevCtx c, int32_t begin, int32_t end, int merged) {
    extern __shared__ double smem_d[];
    double* Ds = smem_d;                   // [SBLK*SBLK] transposed inverse of the item's diagonal block
    double* xs = Ds + SBLK * SBLK;         // [SBLK]
    double* red = xs + SBLK;               // [SL_CG][SBLK]
    const int tid = threadIdx.x, r = tid & (SBLK - 1), cg = tid >> 7, lane = tid & 31, wq = (tid >> 5) & 3;
    int32_t* bflag = c.flags + c.ndblk;
    for (int32_t it = begin + blockIdx.x; it < end; it += gridDim.x) {
        const SolveItem I = merged ? c.bwd_seq[it] : c.bwd_items[it];
        const int32_t s = I.sn;
        if (c.skip && c.skip[s]) continue;
        if (merged && I.kind == 2) {
            // merged levels: x of every ancestor is final once all items of the in-launch parent have finished
            if (c.bwd_wait[s] >= 0) wait_count(c.dep_cnt + c.nsuper + c.bwd_wait[s], c.bwd_nitems[c.bwd_wait[s]], c.info);
            // rows below the columns of a tall supernode (what k_bwd_below does in the per-level path): accumulate into bacc,
            // then tell the supernode's block items that one more below item is in place
            bwd_below_item(c, s, I.blk, I.r0, I.nr, red);
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(c.dep_cnt + 2 * c.nsuper + s, 1);
            continue;
        }
        const int32_t f = c.sn_first[s];
        const int32_t nc = c.sn_first[s + 1] - f;
        const int64_t rp = c.sn_rowptr[s];
        const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
        const int32_t nrow = ld;
        const double* panel = c.Lx + c.sn_xptr[s];
        const int32_t db = c.sn_dblk[s];
        const int32_t ncb = (nc + SBLK - 1) / SBLK;
        const int32_t i = I.blk;
        {
            const double* src = c.DinvT + (int64_t)(db + i) * SBLK * SBLK;
            for (int e = tid; e < SBLK * SBLK / 2; e += SL_THREADS) cp_async16(Ds + 2 * e, src + 2 * e);
            cp_async_commit();
        }
        // own right-hand side and signs: off the critical chain
        double bown = 0.0, sown = 1.0;
        if (tid < SBLK && tid < I.nr) {
            bown = __ldcg(c.wk + f + i * SBLK + tid);
            sown = (double)c.sign[f + i * SBLK + tid];
        }
        // merged levels: x of every ancestor is final once all items of the in-launch parent have finished, and this
        // supernode's below items (kind 2) have accumulated into bacc -- waited for here, after the prefetch of the inverse
        // diagonal block and the own right-hand side have been issued
        if (merged && c.bwd_wait[s] >= 0) wait_count(c.dep_cnt + c.nsuper + c.bwd_wait[s], c.bwd_nitems[c.bwd_wait[s]], c.info);
        if (merged && c.bwd_nbelow[s] > 0) wait_count(c.dep_cnt + 2 * c.nsuper + s, c.bwd_nbelow[s], c.info);
        // p[cc] accumulates sum_r L[r, i*128 + cg*32 + cc] * x[r] over this thread's rows r (mod 128)
        double p[32];
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) p[cc] = 0.0;
        const double* colbase = panel + (int64_t)(i * SBLK + cg * 32) * ld;
        const int32_t ncv = min(32, max(0, I.nr - cg * 32));     // valid columns of this group
        // rows below the supernode's columns: their x is final (ancestors were solved earlier).  Tall supernodes had
        // this part accumulated into bacc by k_bwd_below (consumed and re-zeroed here).
        const bool split = c.sn_split[s] != 0;
        double pown = 0.0;
        if (split && tid < SBLK && tid < I.nr) {
            pown = __ldcg(c.bacc + f + i * SBLK + tid);
            c.bacc[f + i * SBLK + tid] = 0.0;
        }
        for (int32_t r0 = split ? nrow : nc; r0 < nrow; r0 += SBLK) {
            const int32_t rr = r0 + r;
            if (rr < nrow) {
                const double xr = __ldcg(c.wk + c.sn_rows[rp + rr]);
#pragma unroll
                for (int cc = 0; cc < 32; ++cc)
                    if (cc < ncv) p[cc] += colbase[(int64_t)cc * ld + rr] * xr;
            }
        }
        // later diagonal blocks except the adjacent one: their x has been available for a while
        for (int32_t j = ncb - 1; j > i + 1; --j) {
            wait_flag(bflag + db + j, c.info);
            const int32_t rr = j * SBLK + r;
            if (rr < nc) {
                const double xr = __ldcg(c.wk + f + rr);
#pragma unroll
                for (int cc = 0; cc < 32; ++cc)
                    if (cc < ncv) p[cc] += colbase[(int64_t)cc * ld + rr] * xr;
            }
        }
        // reduce over the 32 rows of the warp: lane l ends with the sum of column l of its group
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < off; ++k) {
                const double send = up ? p[k] : p[k + off];
                const double keep = up ? p[k + off] : p[k];
                p[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        __syncthreads();
        red[wq * SBLK + cg * 32 + lane] = p[0];          // 4 warps (row quarters) per column group
        __syncthreads();
        double part = 0.0;                               // thread tid < 128 owns column tid of block i
        if (tid < SBLK) part = pown + red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid];
        // the adjacent block (i+1) is the one on the critical chain: its tile was stored transposed at
        // factorisation time, so it can be prefetched into registers before the wait and applied row-wise
        if (i + 1 < ncb) {
            const double* T = c.LsubT + (int64_t)(db + i) * SBLK * SBLK;
            double t[32];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) t[cc] = T[(cg * 32 + cc) * SBLK + r];
            wait_flag(bflag + db + i + 1, c.info);
            if (tid < SBLK) xs[tid] = ((i + 1) * SBLK + tid < nc) ? __ldcg(c.wk + f + (i + 1) * SBLK + tid) : 0.0;
            __syncthreads();
            double a1 = 0.0;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) a1 += t[cc] * xs[cg * 32 + cc];
            __syncthreads();
            red[cg * SBLK + r] = a1;
            __syncthreads();
            if (tid < SBLK) part += red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid];
            __syncthreads();
        }
        if (tid < SBLK) xs[tid] = (tid < I.nr) ? (sown * bown - part) : 0.0;
        cp_async_wait_all();
        __syncthreads();
        // x = X' * tmp : row r of X' is column r of X; Ds holds X' column-major -> Ds[col*128 + row]
        double a2 = 0.0;
#pragma unroll
        for (int cc = 0; cc < 32; ++cc) a2 += Ds[(cg * 32 + cc) * SBLK + r] * xs[cg * 32 + cc];
        __syncthreads();
        red[cg * SBLK + r] = a2;
        __syncthreads();
        if (tid < SBLK && tid < I.nr)
            __stcg(c.wk + f + i * SBLK + tid, red[tid] + red[SBLK + tid] + red[2 * SBLK + tid] + red[3 * SBLK + tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            st_release(bflag + db + i, 1);
            if (merged) atomicAdd(c.dep_cnt + c.nsuper + s, 1);
        }
    }
}

// backward sweep, rows below the columns of a tall supernode: bacc[column] += sum_rows L[row, column] x[row] for one
// (column block, row range) item.  Runs as its own launch before the level's block solves: every x[row] is final
// (ancestors), so the items are independent and fill the machine instead of one CTA walking all the rows.
__global__ void __launch_bounds__(SL_THREADS, 1) k_bwd_below(DevCtx c, int32_t begin) {
    __shared__ double red[SL_CG * SBLK];
    const BelowItem I = c.bwd_below[begin + blockIdx.x];
    if (c.skip && c.skip[I.sn]) return;
    bwd_below_item(c, I.sn, I.blk, I.r0, I.nr, red);
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

// dx_j = d_j (sum_p A[p, j] dy[row_p] - xi_d_j) with one thread per column; a column longer than LONG_COL entries (the dense
// columns of BASELINE config 5 hold 25 000) is left to k_k1_recover_long, one CTA per column (a single lane walking it was the
// whole cost of rhs + recovery on config 5: 4.8 ms; a warp per column still serialised the eight adjacent dense columns: 3.1 ms)
__global__ void k_k1_recover(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ xi_d,
                             double* __restrict__ dx, double* __restrict__ dy) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < A.n) {
        const int64_t b = A.colptr[j], e = A.colptr[j + 1];
        if (e - b <= LONG_COL) {
            double v = 0.0;
            for (int64_t p = b; p < e; ++p) v += A.val[p] * c.wk[c.iperm[A.rowidx[p]]];
            dx[j] = d[j] * (v - xi_d[j]);
        }
    }
    if (j < A.m) dy[j] = c.wk[c.iperm[j]];
}
__global__ void __launch_bounds__(256) k_k1_recover_long(DevCtx c, DevMat A, const double* __restrict__ d, const double* __restrict__ xi_d,
                                                        double* __restrict__ dx) {
    __shared__ double red[8];
    const int64_t j = A.long_cols[blockIdx.x];
    double v = 0.0;
    for (int64_t p = A.colptr[j] + threadIdx.x; p < A.colptr[j + 1]; p += blockDim.x) v += A.val[p] * c.wk[c.iperm[A.rowidx[p]]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        dx[j] = d[j] * (t - xi_d[j]);
    }
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

__global__ void k_zero_unowned(DevCtx c, const int8_t* __restrict__ keep) {
    const int32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < c.N && !keep[q]) c.wk[q] = 0.0;
}

// multi-GPU: the separator ("top") entries of the work vector, packed for the small all-reduce between the local forward
// sweeps and the replicated top solve (a 512-vector per solve on config 4), and unpacked afterwards
__global__ void k_gather_top(const double* __restrict__ wk, const int32_t* __restrict__ cols, int32_t ntop, double* __restrict__ buf) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntop) buf[i] = wk[cols[i]];
}
__global__ void k_scatter_top(double* __restrict__ wk, const int32_t* __restrict__ cols, int32_t ntop, const double* __restrict__ buf) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntop) wk[cols[i]] = buf[i];
}
// multi-GPU: the factorisation status words, arranged so that ONE max-all-reduce makes them identical on every rank
// (info[0] = smallest bad pivot column -> INT_MAX - info[0]; info[2], info[3] = time-out flags), and back
__global__ void k_pack_info(const int32_t* __restrict__ info, int32_t* __restrict__ tmp) {
    if (threadIdx.x == 0) { tmp[0] = 0x7fffffff - info[0]; tmp[1] = info[2]; tmp[2] = info[3]; tmp[3] = 0; }
}
__global__ void k_unpack_info(int32_t* __restrict__ info, const int32_t* __restrict__ tmp) {
    if (threadIdx.x == 0) { info[0] = 0x7fffffff - tmp[0]; info[2] = tmp[1]; info[3] = tmp[2]; }
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

static constexpr size_t INV_SMEM = (size_t)SBLK * (SBLK + 1) * 8;
static constexpr size_t INV2_SMEM = ((size_t)SBLK * LDI2 + 64 * LDT2 + SBLK) * 8;
static constexpr size_t SL_SMEM = ((size_t)SBLK * SBLK + SBLK + (size_t)SL_CG * SBLK) * 8;

#define SETATTR(k, bytes)                                                                                   \
    do {                                                                                                    \
        cudaError_t e_ = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
        if (e_ != cudaSuccess) return e_;                                                                   \
    } while (0)

cudaError_t kernels_static_init() {
    SETATTR(k_small_factor, 96 * 1024);
    SETATTR(k_invert_diag, INV_SMEM);
    SETATTR(k_invert_diag2, INV2_SMEM);
    SETATTR(k_fwd_large, SL_SMEM);
    SETATTR(k_bwd_large, SL_SMEM);
    return factor_kernels_static_init();
}

void launch_small_factor(const DevCtx& c, int32_t begin, int32_t end, size_t smem, cudaStream_t st) {
    if (end > begin) k_small_factor<<<end - begin, SMALL_THREADS, smem, st>>>(c, begin);
}
// TLPB200_INVERT_KERNEL=0: round-1 k_invert_diag
void launch_invert_diag(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    static const bool old_kernel = [] { const char* e = getenv("TLPB200_INVERT_KERNEL"); return e && atoi(e) == 0; }();
    if (end <= begin) return;
    if (old_kernel) k_invert_diag<<<end - begin, INV_THREADS, INV_SMEM, st>>>(c, begin);
    else k_invert_diag2<<<end - begin, INV2_THREADS, INV2_SMEM, st>>>(c, begin);
}
void launch_fwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_fwd_small<<<nblk(end - begin, SOLVE_SMALL_WARPS), 32 * SOLVE_SMALL_WARPS, 0, st>>>(c, begin, end);
}
void launch_bwd_small(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_bwd_small<<<nblk(end - begin, SOLVE_SMALL_WARPS), 32 * SOLVE_SMALL_WARPS, 0, st>>>(c, begin, end);
}
void launch_fwd_large(const DevCtx& c, int32_t begin, int32_t end, int nsm, int merged, cudaStream_t st) {
    if (end > begin) k_fwd_large<<<min(end - begin, nsm), SL_THREADS, SL_SMEM, st>>>(c, begin, end, merged);
}
void launch_bwd_large(const DevCtx& c, int32_t begin, int32_t end, int nsm, int merged, cudaStream_t st) {
    if (end > begin) k_bwd_large<<<min(end - begin, nsm), SL_THREADS, SL_SMEM, st>>>(c, begin, end, merged);
}
void launch_bwd_below(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_bwd_below<<<(unsigned)(end - begin), SL_THREADS, 0, st>>>(c, begin);
}
void launch_gather_top(const double* wk, const int32_t* cols, int32_t ntop, double* buf, cudaStream_t st) {
    if (ntop > 0) k_gather_top<<<(unsigned)((ntop + 255) / 256), 256, 0, st>>>(wk, cols, ntop, buf);
}
void launch_scatter_top(double* wk, const int32_t* cols, int32_t ntop, const double* buf, cudaStream_t st) {
    if (ntop > 0) k_scatter_top<<<(unsigned)((ntop + 255) / 256), 256, 0, st>>>(wk, cols, ntop, buf);
}
void launch_pack_info(const int32_t* info, int32_t* tmp, cudaStream_t st) { k_pack_info<<<1, 32, 0, st>>>(info, tmp); }
void launch_unpack_info(int32_t* info, const int32_t* tmp, cudaStream_t st) { k_unpack_info<<<1, 32, 0, st>>>(info, tmp); }
void launch_zero_unowned(const DevCtx& c, const int8_t* keep, cudaStream_t st) {
    if (c.N > 0) k_zero_unowned<<<nblk(c.N, 256), 256, 0, st>>>(c, keep);
}
void launch_k1_rhs(const DevCtx& c, const DevMat& A, const double* d, const double* xi_p, const double* xi_d,
                   cudaStream_t st) {
    if (A.m > 0) k_k1_rhs<<<nblk(A.m, 128), 128, 0, st>>>(c, A, d, xi_p, xi_d);
}
void launch_k1_recover(const DevCtx& c, const DevMat& A, const double* d, const double* xi_d, double* dx, double* dy,
                       cudaStream_t st) {
    const int64_t mx = A.n > A.m ? A.n : A.m;
    if (mx > 0) k_k1_recover<<<nblk(mx, 128), 128, 0, st>>>(c, A, d, xi_d, dx, dy);
    if (A.nlong > 0) k_k1_recover_long<<<A.nlong, 256, 0, st>>>(c, A, d, xi_d, dx);
}
void launch_k2_rhs(const DevCtx& c, const DevMat& A, const double* xi_p, const double* xi_d, cudaStream_t st) {
    if (A.n + A.m > 0) k_k2_rhs<<<nblk(A.n + A.m, 256), 256, 0, st>>>(c, A, xi_p, xi_d);
}
void launch_k2_recover(const DevCtx& c, const DevMat& A, double* dx, double* dy, cudaStream_t st) {
    if (A.n + A.m > 0) k_k2_recover<<<nblk(A.n + A.m, 256), 256, 0, st>>>(c, A, dx, dy);
}

}  // namespace tlp
