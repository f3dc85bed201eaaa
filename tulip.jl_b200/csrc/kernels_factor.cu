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

#include <algorithm>
#include <cstdlib>

namespace tlp {

// scripts/chain_bench.cu compiles this file with TLP_CHAIN_CLOCKS to read the phase timeline of one CTA
#ifdef TLP_CHAIN_CLOCKS
__device__ unsigned long long g_ticks[2][80];
#define TLP_TICK(idx)                                                            \
    do {                                                                         \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_ticks[0][idx] = clock64();    \
        if (blockIdx.x == 0 && threadIdx.x == 160) g_ticks[1][idx] = clock64();  \
    } while (0)
#else
#define TLP_TICK(idx)
#endif

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// FP64 operations whose relative order the compiler keeps (volatile): the chain kernels below are bound by the latency of
// dependent DFMAs (~18 cycles), so the order in which independent ones are issued between them is the schedule.
__device__ __forceinline__ double vfma(double a, double b, double c) {
    double d;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
    return d;
}
__device__ __forceinline__ double vmul(double a, double b) {
    double d;
    asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
    return d;
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
    TLP_TICK(0);
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
    TLP_TICK(72);
    if (tid == 0) trace_mark(c, pc.level, 0, true);
}

// ------------------------------------------------------------------------------------------
// k_diag_factor2 (round 2): same elimination (unscaled columns u_ij, signed pivots d_j, W = U D^-1), rebuilt on what
// scripts/chain_bench.cu measures on B200: a dependent DFMA costs 8 cycles but ONE warp also issues at most one FP64
// instruction per 8 cycles (a DMMA per 16, 26 dependent), MUFU.RCP64H 19, a shuffle 26, a full 1.0/d 83, and one SM pulls
// only ~16 B/clk out of L2 with 8-byte loads.  k_diag_factor (104 k cycles) is bound by barriers and FMA-issue; here
//  * the panel (8 columns x all rows below) is done by five warps, each holding the eight diagonal rows in lanes 0-7 and 24
//    other rows in lanes 8-31: pivots and the rows of the diagonal block travel by shuffle inside the warp, so there is no
//    shared-memory hand-over and no block-wide barrier inside a panel, and a warp issues 68 FP64 instructions per panel (a
//    private copy of the 8x8 block per thread, the previous layout, cost 220: 1 800 cycles of pure issue).  The reciprocal
//    is the hardware seed plus one cubic step (3 FMAs, error e^3 < 2^-57), with the sign test beside it, off the chain.
//  * the trailing update C -= U W' runs on DMMA 8x8x4 tiles straight out of shared memory, four tiles of a tile row in
//    flight per warp.  The tiles of the NEXT 8 columns are updated first by all warps; then the panel warps eliminate that
//    block while the other eleven finish the rest of the trailing matrix underneath.
//  * block copy in / out with 16 independent loads / stores per thread in flight.
// Layout: Cs column-major [128][130] (130: conflict-free accumulator fragments), U / -W panels [2][8][132] double-buffered.
// ------------------------------------------------------------------------------------------
constexpr int DF2_THREADS = 512;
constexpr int DF2_WARPS = DF2_THREADS / 32;
constexpr int DF2_PANEL_WARPS = 5;                       // 8 diagonal rows (replicated) + 24 rows per warp: 5 x 24 = 120 rows below
constexpr int DF2_TRAIL_WARPS = DF2_WARPS - DF2_PANEL_WARPS;
constexpr int LDC2 = PIECE + 2;
constexpr int LDP2 = PIECE + 4;

// 1 / d to within an ulp or two: the 20-bit hardware seed and one cubic step r0 (1 + e + e^2), e = 1 - d r0  (error e^3)
__device__ __forceinline__ double fast_rcp(double d) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    const double e = fma(-d, r0, 1.0);
    const double p = fma(e, e, e);
    return fma(r0, p, r0);
}

// columns j0 .. j0+7, all rows from j0 down (nrows of them); called by the DF2_PANEL_WARPS panel warps
__device__ __forceinline__ void df2_panel(double* Cs, double* Up, double* Wn, double* dd, double* rdd, const double* sgn,
                                          int32_t gcol0, int32_t* info, int j0, int nb, int nrows, int wp, int lane) {
    const int t = (lane < NBD) ? lane : NBD + 24 * wp + (lane - NBD);      // row offset inside the panel
    const bool live = t < nrows;
    const int i = j0 + (live ? t : 0);
    double u[NBD], an[NBD], sg[NBD];
#pragma unroll
    for (int j = 0; j < NBD; ++j) {
        u[j] = Cs[(j0 + j) * LDC2 + i];
        sg[j] = sgn[j0 + j];
    }
    // warp 0 writes the eliminated diagonal rows back over the block the other panel warps have just read
    asm volatile("bar.sync 1, %0;" ::"n"(DF2_PANEL_WARPS * 32) : "memory");
    unsigned bad = 0;
    double dmine = 1.0, rmine = 1.0;
#pragma unroll
    for (int j = 0; j < NBD; ++j) {
        const double draw = __shfl_sync(0xffffffffu, u[j], j);      // pivot: lane j holds diagonal row j
        const bool ok = draw * sg[j] > 0.0;                           // beside the seed, not in front of it
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(draw));
        r0 = ok ? r0 : sg[j];                                         // failed sign test: pivot replaced by s_j, as in k_diag_factor
        const double d = ok ? draw : sg[j];
        if (!ok) bad |= 1u << j;
        const double e = fma(-d, r0, 1.0);
        const double pe = fma(e, e, e);
        const double r = fma(r0, pe, r0);
        if (lane == j) { dmine = d; rmine = r; }
        const double a = u[j] * r;
        an[j] = a;
#pragma unroll
        for (int k = j + 1; k < NBD; ++k) {
            const double ukj = __shfl_sync(0xffffffffu, u[j], k);     // u_kj from diagonal row k
            u[k] = fma(-a, ukj, u[k]);
        }
    }
    if (!live) return;
    if (lane < NBD) {
        if (wp == 0) {
#pragma unroll
            for (int j = 0; j < NBD; ++j)
                if (j < lane) Cs[(j0 + j) * LDC2 + i] = u[j];
            Cs[(j0 + lane) * LDC2 + i] = dmine;      // the pivot actually used
            dd[j0 + lane] = dmine;
            rdd[j0 + lane] = rmine;
            if (((bad >> lane) & 1u) && lane < nb) atomicMin(info, gcol0 + j0 + lane);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NBD; ++j) {
            Cs[(j0 + j) * LDC2 + i] = u[j];
            Up[j * LDP2 + i] = u[j];
            Wn[j * LDP2 + i] = -an[j];
        }
    }
}

// n <= 4 tiles (ti, tk0 .. tk0+n-1) of one tile row: C[8ti.., 8tk..] += U[8ti.., 0:8] * Wn[8tk.., 0:8]'.  The second k-step of a
// tile issues four DMMAs (64 cycles) after its first: no wait on the 26-cycle accumulate latency.
__device__ __forceinline__ void df2_quad(double* Cs, const double* Up, const double* Wn, int ti, int tk0, int n, int g, int t4) {
    const double a0 = Up[t4 * LDP2 + 8 * ti + g], a1 = Up[(4 + t4) * LDP2 + 8 * ti + g];
    double b0[4], b1[4], c0[4], c1[4];
    double* pc[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int tk = tk0 + min(x, n - 1);      // past the end: the last tile again, result dropped
        b0[x] = Wn[t4 * LDP2 + 8 * tk + g];
        b1[x] = Wn[(4 + t4) * LDP2 + 8 * tk + g];
        pc[x] = Cs + (8 * tk + 2 * t4) * LDC2 + 8 * ti + g;
        c0[x] = pc[x][0];
        c1[x] = pc[x][LDC2];
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) dmma884(c0[x], c1[x], a0, b0[x]);
#pragma unroll
    for (int x = 0; x < 4; ++x) dmma884(c0[x], c1[x], a1, b1[x]);
#pragma unroll
    for (int x = 0; x < 4; ++x)
        if (x < n) {
            pc[x][0] = c0[x];
            pc[x][LDC2] = c1[x];
        }
}

// one tile: the strip of the next block (two dependent DMMAs)
__device__ __forceinline__ void df2_tile1(double* Cs, const double* Up, const double* Wn, int ti, int tk, int g, int t4) {
    const double a0 = Up[t4 * LDP2 + 8 * ti + g], a1 = Up[(4 + t4) * LDP2 + 8 * ti + g];
    const double b0 = Wn[t4 * LDP2 + 8 * tk + g], b1 = Wn[(4 + t4) * LDP2 + 8 * tk + g];
    double* pc = Cs + (8 * tk + 2 * t4) * LDC2 + 8 * ti + g;
    double c0 = pc[0], c1 = pc[LDC2];
    dmma884(c0, c1, a0, b0);
    dmma884(c0, c1, a1, b1);
    pc[0] = c0;
    pc[LDC2] = c1;
}

__global__ void __launch_bounds__(DF2_THREADS, 1) k_diag_factor2(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Cs = smem_d;                       // [PIECE][LDC2] column-major
    double* Upan = Cs + PIECE * LDC2;          // [2][NBD][LDP2]  u_ij of the current block
    double* Wpan = Upan + 2 * NBD * LDP2;      // [2][NBD][LDP2]  -u_ij / d_j
    double* dd = Wpan + 2 * NBD * LDP2;        // [PIECE] pivots
    double* rdd = dd + PIECE;                  // [PIECE] reciprocals (later: 1 / (s_j l_jj))
    double* sgn = rdd + PIECE;                 // [PIECE] expected pivot signs
    const Piece pc = c.pieces[c.level_pieces[begin + blockIdx.x]];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t lc0 = pc.c0 - f, w = pc.c1 - pc.c0;
    double* D = c.Lx + c.sn_xptr[s] + (int64_t)lc0 * ld + lc0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int il = tid & (PIECE - 1), q4 = tid >> 7;      // 4 columns per pass of the block copy
    const int nt = (w + NBD - 1) / NBD, nt8 = nt * NBD;

    if (tid == 0) trace_mark(c, pc.level, 0, false);
    TLP_TICK(0);
    // lower triangle in, 16 loads in flight per thread; zero elsewhere in the nt8 x 128 window except above the diagonal
    // tiles, which nothing reads
    if (tid < PIECE) sgn[tid] = (tid < w) ? (double)c.sign[pc.c0 + tid] : 1.0;
    for (int kb = 0; kb < nt8; kb += 64) {
        double v[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = kb + q4 + 4 * x;
            v[x] = (k < w && il < w && il >= k) ? D[(int64_t)k * ld + il] : 0.0;
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = kb + q4 + 4 * x;
            if (k < nt8 && (il | 31) >= (k & ~7)) Cs[k * LDC2 + il] = v[x];
        }
    }
    __syncthreads();
    if (tid >= w && tid < nt8) Cs[tid * LDC2 + tid] = 1.0;      // unit diagonal on the padding columns
    __syncthreads();
    TLP_TICK(1);
    if (warp < DF2_PANEL_WARPS) df2_panel(Cs, Upan, Wpan, dd, rdd, sgn, pc.c0, c.info, 0, min(NBD, w), nt8, warp, lane);
    TLP_TICK(2);
    __syncthreads();
    for (int b = 0; b + 1 < nt; ++b) {
        const double* Up = Upan + (b & 1) * NBD * LDP2;
        const double* Wn = Wpan + (b & 1) * NBD * LDP2;
        const int t1 = b + 1;               // first trailing tile row / column
        const int nrem = nt - t1;           // tile rows left (<= 15)
        // (1) columns of the next block: tiles (t1 + x, t1), one per warp
        if (warp < nrem) df2_tile1(Cs, Up, Wn, t1 + warp, t1, g, t4);
        TLP_TICK(3 + 4 * b);
        __syncthreads();
        TLP_TICK(4 + 4 * b);
        // (2) panel warps eliminate the next block; the others finish the trailing matrix (tile columns >= t1 + 1)
        if (warp < DF2_PANEL_WARPS) {
            const int j0 = t1 * NBD;
            df2_panel(Cs, Upan + (t1 & 1) * NBD * LDP2, Wpan + (t1 & 1) * NBD * LDP2, dd, rdd, sgn, pc.c0, c.info, j0, min(NBD, w - j0), nt8 - j0,
                      warp, lane);
        } else {
            // trailing tile rows p = 0 .. T-1 (row p: tiles q = 0 .. p), cut into quads (p, 4qq .. 4qq+3), p >= 4qq, listed column
            // quad by column quad; quad number idx -> warp idx mod DF2_TRAIL_WARPS
            const int T = nrem - 1, c0t = t1 + 1;
            int nq = 0;
            for (int qq = 0; 4 * qq < T; ++qq) nq += T - 4 * qq;
            for (int idx = warp - DF2_PANEL_WARPS; idx < nq; idx += DF2_TRAIL_WARPS) {
                int rem = idx, qq = 0;
                while (rem >= T - 4 * qq) { rem -= T - 4 * qq; ++qq; }
                const int p = 4 * qq + rem;
                df2_quad(Cs, Up, Wn, c0t + p, c0t + 4 * qq, min(4, p + 1 - 4 * qq), g, t4);
            }
        }
        TLP_TICK(5 + 4 * b);
        __syncthreads();
        TLP_TICK(6 + 4 * b);
    }
    // l_jj = sqrt(|d_j|), l_ij = u_ij / (s_j l_jj)
    TLP_TICK(70);
    if (tid < w) {
        const double sk = sgn[tid];
        const double l = sqrt(dd[tid] * sk);
        dd[tid] = l;
        rdd[tid] = 1.0 / (sk * l);
    }
    __syncthreads();
    TLP_TICK(71);
    for (int kb = 0; kb < w; kb += 64) {
        double v[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = kb + q4 + 4 * x;
            v[x] = (k < w && il < w && il >= k) ? ((il == k) ? dd[k] : Cs[k * LDC2 + il] * rdd[k]) : 0.0;
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = kb + q4 + 4 * x;
            if (k < w && il < w && il >= k) D[(int64_t)k * ld + il] = v[x];
        }
    }
    TLP_TICK(72);
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
    TLP_TICK(0);
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
    TLP_TICK(20);
    if (tid == 0) trace_mark(c, pc.level, 1, true);
}

// ------------------------------------------------------------------------------------------
// k_trsm2 (round 2): the same solve with every global load issued once, up front, as asynchronous copies: L11 (k-major, as
// stored) and the 64 x w block of rows both sit in shared memory, so the eight 16-column steps run without an L2 round trip
// each (k_trsm fetched its block of L11, the diagonal block and the accumulators from global memory inside every step;
// scripts/chain_bench.cu: 50.8 k cycles per CTA).  64 rows per CTA, two CTAs per 128-row panel task.  Solved columns are
// kept in shared memory as -s_k x_k, so the products X[:, 0:j0] S L11[j0:j0+16, 0:j0]' need no operand fix-up and use two
// accumulator sets (even / odd k-steps: a dependent DMMA costs ~100 cycles).  The 16 x 16 diagonal blocks are pre-scaled
// (c_jk = L_jk s_k / (s_k L_kk)) so that the substitution chain is one FMA per column instead of a multiply and an FMA.
// ------------------------------------------------------------------------------------------
constexpr int TR2_THREADS = 256;
constexpr int TR2_ROWS = 64;
constexpr int LDB2 = PIECE + 4;      // 132
constexpr int LDX2 = TR2_ROWS + 4;   // 68
constexpr int LDCD = TRB + 1;        // 17

__global__ void __launch_bounds__(TR2_THREADS, 1) k_trsm2(DevCtx c, int32_t begin) {
    extern __shared__ double smem_d[];
    double* Bs = smem_d;                       // [PIECE][LDB2]  Bs[k][n] = L11[n, k]  (n >= k, else 0)
    double* Xs = Bs + PIECE * LDB2;            // [PIECE][LDX2]  Xs[k][r]: rows of the panel; solved columns as -s_k x_k
    double* Cd = Xs + PIECE * LDX2;            // [PIECE][LDCD]  Cd[k][j] = L11[j0 + j, k] / L11[k, k] inside k's 16-block (j0 + j > k)
    double* invd = Cd + PIECE * LDCD;          // [PIECE]        1 / (s_k L11[k, k])
    double* sgd = invd + PIECE;                // [PIECE]        s_k
    const PanelTask T = c.panel[begin + (blockIdx.x >> 1)];
    const int half = blockIdx.x & 1;
    const int32_t nr = min(TR2_ROWS, T.nr - half * TR2_ROWS);
    if (nr <= 0) return;
    const int32_t r0 = T.r0 + half * TR2_ROWS;
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const int32_t kb = pc.c0 - f, w = pc.c1 - pc.c0;
    double* X = c.Lx + c.sn_xptr[s];
    const double* L11 = X + (int64_t)kb * ld + kb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    if (tid == 0 && half == 0) trace_mark(c, pc.level, 1, false);
    TLP_TICK(0);
    const int nblk = (w + TRB - 1) / TRB, w16 = nblk * TRB, ngrp = (w16 + 31) / 32;

    // Operands arrive in groups of 32 columns; the loads of group g + 1 are in flight (in registers) while the two 16-column
    // steps of group g run: one SM pulls only ~16 B/clk out of L2, all 130 KB up front cost 10 k cycles of the 38 k.
    if (tid < PIECE) sgd[tid] = (tid < w) ? (double)c.sign[pc.c0 + tid] : 1.0;
    const int n = tid & (PIECE - 1), kq = tid >> 7;          // L11: row n, columns 32 g + kq + 2 x
    const int r = tid & (TR2_ROWS - 1), kq4 = tid >> 6;      // rows of the panel: row r, columns 32 g + kq4 + 4 x
    double vb[16], vx[8];
    auto issue = [&](int gidx) {
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = 32 * gidx + kq + 2 * x;
            vb[x] = (k < w && n < w && n >= k) ? L11[(int64_t)k * ld + n] : 0.0;
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int k = 32 * gidx + kq4 + 4 * x;
            vx[x] = (k < w && r < nr) ? X[(int64_t)(kb + k) * ld + r0 + r] : 0.0;
        }
    };
    auto commit = [&](int gidx) {
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const int k = 32 * gidx + kq + 2 * x;
            if (k < w16 && n >= 32 * gidx) Bs[k * LDB2 + n] = vb[x];      // rows above the group's diagonal blocks: never read
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int k = 32 * gidx + kq4 + 4 * x;
            if (k < w16) Xs[k * LDX2 + r] = vx[x];
        }
    };
    issue(0);
    commit(0);
    __syncthreads();
    TLP_TICK(1);

    for (int gidx = 0; gidx < ngrp; ++gidx) {
        if (gidx + 1 < ngrp) issue(gidx + 1);
        if (tid < 32) {
            const int k = 32 * gidx + tid;
            if (k < w16) invd[k] = (k < w) ? 1.0 / (sgd[k] * Bs[k * LDB2 + k]) : 1.0;
        }
        __syncthreads();
        for (int e = tid; e < 32 * TRB; e += TR2_THREADS) {
            const int k = 32 * gidx + (e >> 4), j = e & 15, nn = (k & ~(TRB - 1)) + j;
            if (k < w16) Cd[k * LDCD + j] = (nn > k) ? Bs[k * LDB2 + nn] * sgd[k] * invd[k] : 0.0;
        }
        __syncthreads();
        for (int b = 2 * gidx; b < min(2 * gidx + 2, nblk); ++b) {
            const int j0 = b * TRB;
            if (j0 > 0) {
                // Xs[j0 .. j0+15][rows of this warp] += (-X S)[:, 0 .. j0) * L11[j0 .. j0+15, 0 .. j0)'
                double acc[2][2], acc2[2][2];
                const int rr = warp * 8 + g;
#pragma unroll
                for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        acc[nj][e] = Xs[(j0 + nj * 8 + t4 * 2 + e) * LDX2 + rr];
                        acc2[nj][e] = 0.0;
                    }
#pragma unroll 2
                for (int k4 = 0; k4 < j0; k4 += 8) {
                    const double a0 = Xs[(k4 + t4) * LDX2 + rr], a1 = Xs[(k4 + 4 + t4) * LDX2 + rr];
                    const double b00 = Bs[(k4 + t4) * LDB2 + j0 + g], b01 = Bs[(k4 + t4) * LDB2 + j0 + 8 + g];
                    const double b10 = Bs[(k4 + 4 + t4) * LDB2 + j0 + g], b11 = Bs[(k4 + 4 + t4) * LDB2 + j0 + 8 + g];
                    dmma884(acc[0][0], acc[0][1], a0, b00);
                    dmma884(acc[1][0], acc[1][1], a0, b01);
                    dmma884(acc2[0][0], acc2[0][1], a1, b10);
                    dmma884(acc2[1][0], acc2[1][1], a1, b11);
                }
#pragma unroll
                for (int nj = 0; nj < 2; ++nj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) Xs[(j0 + nj * 8 + t4 * 2 + e) * LDX2 + rr] = acc[nj][e] + acc2[nj][e];
                __syncthreads();
            }
            TLP_TICK(2 + 2 * b);
            {
                // 16 x 16 diagonal block: four threads per row (row = tid / 4, columns j = 4 jj + q); the finished z_k goes to
                // the other lanes by a shuffle: 15 steps of (shuffle + FMA, 34 cycles).  One thread per row issues 120 FMAs at
                // one per 8 cycles, and the compiler orders them as fifteen back-to-back accumulation chains (2 200 cycles).
                const int row = tid >> 2, q = tid & 3, base = lane & ~3;
                double z[4], cf[TRB - 1][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) z[jj] = Xs[(j0 + 4 * jj + q) * LDX2 + row];
#pragma unroll
                for (int k = 0; k < TRB - 1; ++k)
#pragma unroll
                    for (int jj = k >> 2; jj < 4; ++jj) cf[k][jj] = Cd[(j0 + k) * LDCD + 4 * jj + q];      // all up front, off the chain
#pragma unroll
                for (int k = 0; k < TRB - 1; ++k) {
                    const double nz = -__shfl_sync(0xffffffffu, z[k >> 2], base | (k & 3));
#pragma unroll
                    for (int jj = k >> 2; jj < 4; ++jj)
                        if (4 * jj > k || q > k - 4 * jj) z[jj] = fma(nz, cf[k][jj], z[jj]);
                }
                const bool valid = row < nr;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + 4 * jj + q;
                    const double x = z[jj] * invd[j];
                    Xs[j * LDX2 + row] = -sgd[j] * x;
                    if (valid && j < w) X[(int64_t)(kb + j) * ld + r0 + row] = x;
                }
            }
            __syncthreads();
            TLP_TICK(3 + 2 * b);
        }
        if (gidx + 1 < ngrp) {
            commit(gidx + 1);
            __syncthreads();
        }
    }
    if (tid == 0 && half == 0) trace_mark(c, pc.level, 1, true);
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

// kpart / nparts: this CTA multiplies only its share of the piece's columns (critical tiles are split over several CTAs:
// a 64x64x128 tile is 14 k cycles on one SM, and the next level's diagonal block waits for it); nparts > 1 needs atomic = 1
__device__ __forceinline__ void update_tile(const DevCtx& c, const UpdTask& T, double* smem_d, UpdShared& sh, int atomic, int cls,
                                            int kpart = 0, int nparts = 1) {
    const Piece pc = c.pieces[T.piece];
    const int32_t s = pc.sn;
    if (c.skip && c.skip[s]) return;     // uniform per CTA
    if (threadIdx.x == 0) trace_mark(c, pc.level, cls, false);
    TLP_TICK(0);
    const int32_t f = c.sn_first[s];
    const int64_t rp = c.sn_rowptr[s];
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - rp);
    const int32_t* rows = c.sn_rows + rp;
    const int32_t kall = pc.c1 - pc.c0;
    const int32_t kshare = ((kall + nparts - 1) / nparts + KC - 1) / KC * KC;      // whole chunks of KC columns per part
    const int32_t kfirst = kpart * kshare;
    if (kfirst >= kall) return;      // uniform per CTA
    const int32_t kdim = min(kshare, kall - kfirst);
    const double* panel = c.Lx + c.sn_xptr[s] + (int64_t)(pc.c0 - f + kfirst) * ld;
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
    if (tid < kdim) sh.sgk[tid] = (double)c.sign[pc.c0 + kfirst + tid];

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const bool neg = c.has_neg != 0;
    for (int ch = 0; ch < nch; ++ch) {
        if (ch == 1) TLP_TICK(1);
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
    TLP_TICK(2);
    if (threadIdx.x == 0) trace_mark(c, pc.level, cls, true);
}

__global__ void __launch_bounds__(UPD_THREADS, 4) k_update(DevCtx c, int32_t begin, int atomic, int nparts) {
    extern __shared__ double smem_d[];
    __shared__ UpdShared sh;
    const UpdTask T = c.upd[begin + blockIdx.x / nparts];
    update_tile(c, T, smem_d, sh, atomic, 2, blockIdx.x % nparts, nparts);
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

static constexpr size_t DF2_SMEM = ((size_t)PIECE * LDC2 + 4 * (size_t)NBD * LDP2 + 3 * PIECE) * 8;
static constexpr size_t TR2_SMEM = ((size_t)PIECE * LDB2 + (size_t)PIECE * LDX2 + (size_t)PIECE * LDCD + 2 * PIECE) * 8;

// TLPB200_CHAIN_KERNELS: 0 = round-1 k_diag_factor / k_trsm, 1 (default) = k_diag_factor2 / k_trsm2
static int chain_variant() {
    static const int v = [] {
        const char* e = getenv("TLPB200_CHAIN_KERNELS");
        return e ? atoi(e) : 1;
    }();
    return v;
}

cudaError_t factor_kernels_static_init() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_diag_factor2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DF2_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_trsm2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR2_SMEM);
    if (e != cudaSuccess) return e;
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
    if (end <= begin) return;
    const int v = chain_variant();
    if (v >= 1) k_diag_factor2<<<end - begin, DF2_THREADS, DF2_SMEM, st>>>(c, begin);
    else k_diag_factor<<<end - begin, DF_THREADS, DF_SMEM, st>>>(c, begin);
}
void launch_trsm(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end <= begin) return;
    if (chain_variant() >= 1) k_trsm2<<<2 * (end - begin), TR2_THREADS, TR2_SMEM, st>>>(c, begin);
    else k_trsm<<<end - begin, TR_THREADS, TR_SMEM, st>>>(c, begin);
}
// ksplit > 1 (critical tiles, atomic accumulation): each tile's K range is shared by ksplit CTAs
void launch_update(const DevCtx& c, int32_t begin, int32_t end, int atomic, cudaStream_t st, int ksplit) {
    if (end <= begin) return;
    const int np = atomic ? std::max(1, ksplit) : 1;
    k_update<<<(end - begin) * np, UPD_THREADS, UPD_SMEM, st>>>(c, begin, atomic, np);
}
// persistent work-queue launch: `nsm` SMs x 4 resident CTAs; CTAs on SMs >= nsm - reserve exit immediately
void launch_update_lazy(const DevCtx& c, int32_t begin, int32_t end, int32_t* counter, int nsm, int reserve, cudaStream_t st) {
    if (end <= begin) return;
    const int grid = min(end - begin + 4 * reserve, 4 * nsm);
    k_update_lazy<<<grid, UPD_THREADS, UPD_SMEM, st>>>(c, begin, end, counter, nsm - reserve);
}

}  // namespace tlp
