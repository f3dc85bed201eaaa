// Dense kernels of the numeric factorisation of non-small supernodes (sm_100a), processed in
// column pieces of <= 128 columns:
//   k_diag_factor : signed Cholesky of the piece's w x w diagonal block        (one CTA, shared memory)
//   k_trsm        : X = A21 * L11^{-T} * S for the rows below, 128 rows per CTA (FP64 DMMA + substitution)
//   k_update      : C_tgt -= L[I,piece] * S * L[K,piece]'  64x64 tiles          (FP64 DMMA, cp.async pipeline)
// Together they are what cholesky!(F, Symmetric(K)) / ldlt!(F, ...) do per supernode in the
// reference's CHOLMOD backend (/root/reference/src/KKT/Cholmod/spd.jl:46, sqd.jl:53) and what
// src/KKT/Dense/lapack.jl:95 does on the whole matrix.
//
// FP64 on B200: DMMA.8x8x4 and DFMA both peak at 64 FMA/clk/SM (measured 37.1 TFLOP/s,
// scripts/dmma_bench.cu); every mma.sync f64 shape lowers to DMMA.8x8x4.  tcgen05 has no FP64 kind.
#include "kernels.cuh"

namespace tlp {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit_() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait_() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ int32_t pos_in_target_(const DevCtx& c, int32_t t, int32_t gi) {
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

// ------------------------------------------------------------------------------------------
// k_diag_factor: blocked right-looking LDL'-style elimination (columns kept unscaled,
// u_ij = l_ij * s_j * l_jj, pivots d_j) in blocks of 8 columns with one-block look-ahead:
// warp 0 eliminates the next 8x8 diagonal block in registers (shuffles) while the other 15 warps
// apply the rank-8 update of the current block to the rest of the trailing matrix.
// ------------------------------------------------------------------------------------------
constexpr int DF_THREADS = 512;
constexpr int LDD = PIECE + 1;
constexpr int NBD = 8;
constexpr int LDW = NBD + 1;

__device__ __forceinline__ void diag_block_warp0(double* Cs, double* dd, double* rdd, double* Wd, const double* sgn, int32_t gcol0,
                                                  int32_t* info, int j0, int nb, int lane) {
    double u[NBD];
#pragma unroll
    for (int k = 0; k < NBD; ++k) u[k] = (lane < nb && k <= lane) ? Cs[(j0 + k) * LDD + j0 + lane] : 0.0;
    double dmine = 1.0;
#pragma unroll
    for (int j = 0; j < NBD; ++j) {
        double d = __shfl_sync(0xffffffffu, u[j], j);
        if (j < nb) {
            const double sj = sgn[j0 + j];
            if (!(d * sj > 0.0)) {
                if (lane == 0) atomicMin(info, gcol0 + j0 + j);
                d = sj;
            }
            if (lane == j) dmine = d;
            const double a = u[j] * (1.0 / d);      // one reciprocal per column, shared by all lanes
#pragma unroll
            for (int k = j + 1; k < NBD; ++k) {
                const double ukj = __shfl_sync(0xffffffffu, u[j], k);
                if (lane >= k) u[k] -= a * ukj;
            }
        }
    }
    if (lane < nb) { dd[j0 + lane] = dmine; rdd[j0 + lane] = 1.0 / dmine; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NBD; ++k)
        if (lane < nb && k <= lane) {
            Cs[(j0 + k) * LDD + j0 + lane] = u[k];
            if (k < lane) Wd[lane * LDW + k] = u[k] * rdd[j0 + k];
        }
}

__global__ void __launch_bounds__(DF_THREADS, 1) k_diag_factor(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Cs = smem_d;                 // [PIECE][LDD]  column-major
    double* dd = Cs + PIECE * LDD;       // [PIECE] pivots
    double* rdd = dd + PIECE;            // [PIECE] their reciprocals (later: 1 / (s_j l_jj))
    double* Wd = rdd + PIECE;            // [NBD][LDW]    u_jj' / d_j' of the current diagonal block
    double* Wp = Wd + NBD * LDW;         // [PIECE][LDW]  u_kj / d_j of the current block, rows below it
    const Piece pc = c.pieces[c.level_pieces[begin + blockIdx.x]];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t lc0 = pc.c0 - f, w = pc.c1 - pc.c0;
    double* D = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0;
    double* sgn = Wp + PIECE * LDW;      // [PIECE] expected pivot signs (shared: they sit on the critical chain)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int il = tid & (PIECE - 1), q4 = tid >> 7;

    if (tid == 0) trace_mark(c, pc.level, 0, false);
    if (tid < w) sgn[tid] = (double)c.sign[pc.c0 + tid];
    for (int kb = 0; kb < w; kb += 32) {       // 8 independent loads in flight per thread
        double v[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int k = kb + q4 + 4 * x;
            v[x] = (k < w && il < w && il >= k) ? D[(int64_t)k * ld + il] : 0.0;
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int k = kb + q4 + 4 * x;
            if (k < w && il < w) Cs[k * LDD + il] = v[x];
        }
    }
    __syncthreads();
    const int nblk = (w + NBD - 1) / NBD;
    if (warp == 0) diag_block_warp0(Cs, dd, rdd, Wd, sgn, pc.c0, c.info, 0, min(NBD, w), lane);
    __syncthreads();
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * NBD, nb = min(NBD, w - j0), r0 = j0 + nb;
        // (b) rows below the diagonal block: u_ij = a_ij - sum_{j'<j} u_ij' * (u_jj'/d_j')
        if (tid < w - r0) {
            const int i = r0 + tid;
            double u[NBD];
#pragma unroll
            for (int j = 0; j < NBD; ++j) u[j] = (j < nb) ? Cs[(j0 + j) * LDD + i] : 0.0;
#pragma unroll
            for (int j = 1; j < NBD; ++j)
#pragma unroll
                for (int jp = 0; jp < j; ++jp)
                    if (j < nb) u[j] -= u[jp] * Wd[j * LDW + jp];
#pragma unroll
            for (int j = 0; j < NBD; ++j)
                if (j < nb) {
                    Cs[(j0 + j) * LDD + i] = u[j];
                    Wp[i * LDW + j] = u[j] * rdd[j0 + j];
                } else {
                    Wp[i * LDW + j] = 0.0;
                }
        }
        __syncthreads();
        if (r0 >= w) break;
        const int nb1 = min(NBD, w - r0);        // width of the next block
        // (c1) rank-nb update of the next block's columns (all rows >= column)
        {
            const int rows = w - r0;
            for (int e = tid; e < rows * nb1; e += DF_THREADS) {
                const int kk = e / rows, ii = e - kk * rows;
                const int i = r0 + ii, k = r0 + kk;
                if (i >= k) {
                    double a = 0.0;
#pragma unroll
                    for (int j = 0; j < NBD; ++j) a += Cs[(j0 + j) * LDD + i] * Wp[k * LDW + j];
                    Cs[k * LDD + i] -= a;
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            diag_block_warp0(Cs, dd, rdd, Wd, sgn, pc.c0, c.info, r0, nb1, lane);
        } else {
            // (c2) rest of the trailing matrix: columns >= r0 + nb1
            const int c0 = r0 + nb1;
            const int t = tid - 32;                  // 0..479
            const int qq = t / 120, ii = t - qq * 120;
            const int i = c0 + ii;
            if (i < w) {
                double u[NBD];
#pragma unroll
                for (int j = 0; j < NBD; ++j) u[j] = Cs[(j0 + j) * LDD + i];
                // 4 independent accumulators per trip (the 8-term dot products are latency chains)
                for (int k = c0 + qq; k <= i; k += 16) {
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    const int k1 = min(k + 4, w - 1), k2 = min(k + 8, w - 1), k3 = min(k + 12, w - 1);
#pragma unroll
                    for (int j = 0; j < NBD; ++j) {
                        a0 += u[j] * Wp[k * LDW + j];
                        a1 += u[j] * Wp[k1 * LDW + j];
                        a2 += u[j] * Wp[k2 * LDW + j];
                        a3 += u[j] * Wp[k3 * LDW + j];
                    }
                    // every (row i, column) entry has exactly one owner thread; the loads are guarded like the stores so that a
                    // clamped column index (k1..k3 = w - 1 past the end) never reads an entry its owner is writing (racecheck, r2)
                    const double c0v = Cs[k * LDD + i];
                    const double c1v = (k + 4 <= i) ? Cs[(k + 4) * LDD + i] : 0.0;
                    const double c2v = (k + 8 <= i) ? Cs[(k + 8) * LDD + i] : 0.0;
                    const double c3v = (k + 12 <= i) ? Cs[(k + 12) * LDD + i] : 0.0;
                    Cs[k * LDD + i] = c0v - a0;
                    if (k + 4 <= i) Cs[(k + 4) * LDD + i] = c1v - a1;
                    if (k + 8 <= i) Cs[(k + 8) * LDD + i] = c2v - a2;
                    if (k + 12 <= i) Cs[(k + 12) * LDD + i] = c3v - a3;
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // l_jj = sqrt(|d_j|), l_ij = u_ij / (s_j l_jj)
    if (tid < w) {
        const double sk = sgn[tid];
        const double l = sqrt(dd[tid] * sk);
        dd[tid] = l;
        rdd[tid] = 1.0 / (sk * l);
    }
    __syncthreads();
    for (int k = q4; k < w; k += 4)
        if (il < w && il >= k) D[(int64_t)k * ld + il] = (il == k) ? dd[k] : Cs[k * LDD + il] * rdd[k];
    if (tid == 0) trace_mark(c, pc.level, 0, true);
}

// ------------------------------------------------------------------------------------------
// k_trsm: 128 rows per CTA, 8 warps.  Columns are solved in blocks of 16: the contribution of the
// already solved columns is a DMMA product (X[:, 0:j0] * (S L11[jb, 0:j0])'), the 16x16 diagonal
// block is a register substitution (one row per thread).
// ------------------------------------------------------------------------------------------
constexpr int TR_THREADS = 256;
constexpr int TR_ROWS = 128;
constexpr int TRB = 16;
constexpr int LDX = TR_ROWS + 4;   // 132
constexpr int LDLB = TRB + 4;      // 20
constexpr int LDLD = TRB + 1;

__global__ void __launch_bounds__(TR_THREADS, 1) k_trsm(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Xs = smem_d;                       // [PIECE][LDX]   solved columns, k-major
    double* Lb = Xs + PIECE * LDX;             // [PIECE][LDLB]  Lb[k][n] = -L[j0+n, k] s_k   (k < j0)
    double* Tt = Lb + PIECE * LDLB;            // [TRB][LDX]
    double* Ld = Tt + TRB * LDX;               // [TRB][LDLD]    Ld[j][k] = L[j0+j, j0+k] s_{j0+k}
    double* invd = Ld + TRB * LDLD;            // [TRB]
    const PanelTask T = c.panel[begin + blockIdx.x];
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t kb = pc.c0 - f, w = pc.c1 - pc.c0;
    double* X = c.Lx + c.sn_xptr[s];
    const double* L11 = X + (int64_t)kb * ld + kb;
    const int8_t* sgn = c.sign + pc.c0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    if (tid == 0) trace_mark(c, pc.level, 1, false);
    const int nblk = (w + TRB - 1) / TRB;

    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * TRB, nb = min(TRB, w - j0);
        for (int e = tid; e < j0 * TRB; e += TR_THREADS) {
            const int k = e >> 4, n = e & 15;
            Lb[k * LDLB + n] = (n < nb) ? -L11[(int64_t)k * ld + j0 + n] * (double)sgn[k] : 0.0;
        }
        {
            const int j = tid >> 4, k = tid & 15;   // 256 threads = 16 x 16
            Ld[j * LDLD + k] = (j < nb && k < j) ? L11[(int64_t)(j0 + k) * ld + j0 + j] * (double)sgn[j0 + k] : 0.0;
            if (tid < TRB) invd[tid] = (tid < nb) ? 1.0 / ((double)sgn[j0 + tid] * L11[(int64_t)(j0 + tid) * ld + j0 + tid]) : 1.0;
        }
        double acc[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int rr = warp * 16 + mi * 8 + g, cc = nj * 8 + t4 * 2 + e;
                    acc[mi][nj][e] = (rr < T.nr && cc < nb) ? X[(int64_t)(kb + j0 + cc) * ld + T.r0 + rr] : 0.0;
                }
        __syncthreads();
        for (int k4 = 0; k4 < j0; k4 += 4) {
            double a[2], bb[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = Xs[(k4 + t4) * LDX + warp * 16 + mi * 8 + g];
#pragma unroll
            for (int nj = 0; nj < 2; ++nj) bb[nj] = Lb[(k4 + t4) * LDLB + nj * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 2; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], bb[nj]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    Tt[(nj * 8 + t4 * 2 + e) * LDX + warp * 16 + mi * 8 + g] = acc[mi][nj][e];
        __syncthreads();
        if (tid < TR_ROWS) {
            double t[TRB];
#pragma unroll
            for (int j = 0; j < TRB; ++j) t[j] = Tt[j * LDX + tid];
#pragma unroll
            for (int j = 0; j < TRB; ++j) {
                double a = t[j];
#pragma unroll
                for (int k = 0; k < j; ++k) a -= t[k] * Ld[j * LDLD + k];
                t[j] = a * invd[j];
            }
            const bool valid = tid < T.nr;
#pragma unroll
            for (int j = 0; j < TRB; ++j) {
                Xs[(j0 + j) * LDX + tid] = t[j];
                if (valid && j < nb) X[(int64_t)(kb + j0 + j) * ld + T.r0 + tid] = t[j];
            }
        }
        __syncthreads();
    }
    if (tid == 0) trace_mark(c, pc.level, 1, true);
}

// ------------------------------------------------------------------------------------------
// Tile update  C_tgt[pos(I), cols(K)] -= L[I, piece] * S * L[K, piece]'   (supernode SYRK/GEMM + scatter).
// 64x64 output tile per CTA pass, 4 warps (32x32 each = 4x4 DMMA tiles), K = piece width <= 128.
// Operand tiles stream through a 3-stage cp.async ring (8-byte copies: panel columns are only
// 8-byte aligned); the result is reduced into the ancestor panel with RED.ADD.F64.
//   k_update      : one tile per CTA ("urgent" tiles on the main stream).
//   k_update_lazy : persistent work-queue variant for the bulk of the tiles on the side stream.  CTAs
//                   that land on one of the last `reserve` SMs exit at once, so those SMs stay free
//                   for the critical-chain kernels (diag / trsm / urgent tiles) of the main stream.
// ------------------------------------------------------------------------------------------
constexpr int UPD_THREADS = 128;
constexpr int KC = 16;
constexpr int LDT = TILE + 4;      // 68: conflict-free fragment loads (68 mod 16 == 4)
constexpr int UPD_STAGES = 3;
constexpr size_t UPD_SMEM = (size_t)UPD_STAGES * 2 * KC * LDT * 8;

struct UpdShared {
    int32_t tpos[TILE];
    int64_t tcol[TILE];
    double sgk[PIECE];
    int32_t next;
};

__device__ __forceinline__ void update_tile(const DevCtx& c, const UpdTask& T, double* smem_d, UpdShared& sh, int atomic, int cls) {
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;     // uniform per CTA
    if (threadIdx.x == 0) trace_mark(c, pc.level, cls, false);
    const int32_t f = c.sn_first[s];
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    const double* panel = c.Lx + c.sn_xptr[s] + (int64_t)(pc.c0 - f) * ld;
    const int32_t kdim = pc.c1 - pc.c0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 1, wc = warp & 1, g = lane >> 2, t4 = lane & 3;
    const int nch = (kdim + KC - 1) / KC;

    auto As = [&](int st, int k, int r) -> double* { return smem_d + ((size_t)(st * 2 + 0) * KC + k) * LDT + r; };
    auto Bs = [&](int st, int k, int r) -> double* { return smem_d + ((size_t)(st * 2 + 1) * KC + k) * LDT + r; };
    auto issue = [&](int ch, int st) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int e = tid + UPD_THREADS * r;
            const int row = e & (TILE - 1), kl = e >> 6, kk = ch * KC + kl;
            const bool kv = kk < kdim;
            const double* col = panel + (int64_t)(kv ? kk : 0) * ld;
            const bool va = kv && row < T.ni, vb = kv && row < T.nk;
            cp_async8_zfill(As(st, kl, row), col + T.i0 + (va ? row : 0), va);
            cp_async8_zfill(Bs(st, kl, row), col + T.k0 + (vb ? row : 0), vb);
        }
    };
#pragma unroll
    for (int st = 0; st < UPD_STAGES - 1; ++st) {
        if (st < nch) issue(st, st);
        cp_async_commit_();
    }

    // target addressing (overlaps with the first copies)
    const int32_t t = T.tgt;
    const int32_t ft = c.sn_first[t];
    const int64_t ldt = c.sn_rowptr[t + 1] - c.sn_rowptr[t];
    double* Tx = c.Lx + c.sn_xptr[t];
    if (tid < TILE) {
        int32_t p = 0;
        if (tid < T.ni) p = (t == s) ? (T.i0 + tid) : pos_in_target_(c, t, rows[T.i0 + tid]);
        sh.tpos[tid] = p;
    } else {
        const int kk = tid - TILE;
        sh.tcol[kk] = (kk < T.nk) ? (int64_t)(rows[T.k0 + kk] - ft) * ldt : 0;
    }
    if (tid < kdim) sh.sgk[tid] = (double)c.sign[pc.c0 + tid];

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const bool neg = c.has_neg != 0;
    for (int ch = 0; ch < nch; ++ch) {
        cp_async_wait_<UPD_STAGES - 2>();
        __syncthreads();
        const int nx = ch + UPD_STAGES - 1;
        if (nx < nch) issue(nx, nx % UPD_STAGES);
        cp_async_commit_();
        const int st = ch % UPD_STAGES;
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = *As(st, k4 + t4, wr * 32 + mi * 8 + g);
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) b[nj] = *Bs(st, k4 + t4, wc * 32 + nj * 8 + g);
            if (neg) {
                const int kk = ch * KC + k4 + t4;
                const double sk = (kk < kdim) ? sh.sgk[kk] : 1.0;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] *= sk;
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
        }
    }

#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ii = wr * 32 + mi * 8 + g, kk = wc * 32 + nj * 8 + t4 * 2 + e;
                if (ii < T.ni && kk < T.nk && (T.diag == 0 || ii >= kk)) {
                    double* p = Tx + sh.tcol[kk] + sh.tpos[ii];
                    if (atomic) atomicAdd(p, -acc[mi][nj][e]); else *p -= acc[mi][nj][e];
                }
            }
    if (threadIdx.x == 0) trace_mark(c, pc.level, cls, true);
}

__global__ void __launch_bounds__(UPD_THREADS, 4) k_update(DevCtx c, int32_t begin, int atomic) {
    extern __shared__ double smem_d[];
    __shared__ UpdShared sh;
    const UpdTask T = c.upd[begin + blockIdx.x];
    update_tile(c, T, smem_d, sh, atomic, 2);
}

__global__ void __launch_bounds__(UPD_THREADS, 4) k_update_lazy(DevCtx c, int32_t begin, int32_t end, int32_t* counter,
                                                                int32_t first_reserved_sm) {
    extern __shared__ double smem_d[];
    __shared__ UpdShared sh;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    // keep the reserved SMs free for the critical chain -- but never let the last CTA leave
    if ((int32_t)smid >= first_reserved_sm) {
        if (threadIdx.x == 0) sh.next = atomicAdd(counter + 1, 1);
        __syncthreads();
        if (sh.next + 1 < (int32_t)gridDim.x) return;
        __syncthreads();
    }
    for (;;) {
        if (threadIdx.x == 0) sh.next = begin + atomicAdd(counter, 1);
        __syncthreads();
        const int32_t task = sh.next;
        if (task >= end) return;
        const UpdTask T = c.upd_lazy[task];
        update_tile(c, T, smem_d, sh, 1, 3);
        cp_async_wait_<0>();
        __syncthreads();       // shared state is reused by the next tile
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static constexpr size_t DF_SMEM = ((size_t)PIECE * LDD + 3 * PIECE + (size_t)NBD * LDW + (size_t)PIECE * LDW) * 8;
static constexpr size_t TR_SMEM = ((size_t)PIECE * LDX + (size_t)PIECE * LDLB + (size_t)TRB * LDX + (size_t)TRB * LDLD + TRB) * 8;

cudaError_t factor_kernels_static_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_diag_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_update_lazy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
    return e;
}

void launch_diag_factor(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_diag_factor<<<end - begin, DF_THREADS, DF_SMEM, st>>>(c, begin);
}
void launch_trsm(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_trsm<<<end - begin, TR_THREADS, TR_SMEM, st>>>(c, begin);
}
void launch_update(const DevCtx& c, int32_t begin, int32_t end, int atomic, cudaStream_t st) {
    if (end > begin) k_update<<<end - begin, UPD_THREADS, UPD_SMEM, st>>>(c, begin, atomic);
}
// persistent work-queue launch: `nsm` SMs x 4 resident CTAs; CTAs on SMs >= nsm - reserve exit immediately
void launch_update_lazy(const DevCtx& c, int32_t begin, int32_t end, int32_t* counter, int nsm, int reserve, cudaStream_t st) {
    if (end <= begin) return;
    const int grid = min(end - begin + 4 * reserve, 4 * nsm);
    k_update_lazy<<<grid, UPD_THREADS, UPD_SMEM, st>>>(c, begin, end, counter, nsm - reserve);
}

}  // namespace tlp
