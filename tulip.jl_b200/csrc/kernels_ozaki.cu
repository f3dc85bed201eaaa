// FP64-accurate supernode Schur updates on the 5th-generation tensor cores (sm_100a):
//     C[I, J] -= L[I, 0:K] * L[J, 0:K]'                     (cholesky! in /root/reference/src/KKT/Cholmod/spd.jl:46;
//                                                            the SYRK of src/KKT/Dense/lapack.jl:85-88 per supernode)
// tcgen05 has no FP64 kind, so the contraction is computed *exactly* on the int8 tensor-core path (Ozaki splitting,
// Ootomo/Ozaki/Yokota 2024) and rounded once:
//
//   * every row r of L has a fixed binary exponent E_r with |L_rk| <= sqrt(K_rr) < 2^E_r  (K = L L', SPD), taken from the
//     row's current diagonal when the supernode starts.  x 2^-E_r is cut into S = 8 signed digits, base 256 after the first
//         x 2^-E = d_0 2^-6 + sum_{p>=1} d_p 2^(-6-8p),   d_0 in [-64, 64], d_p in [-128, 127]     (k_oz_slice, exact, 63 bits)
//     stored as int8 planes in the layout the tensor core reads (K-major, no swizzle, 8x16-byte core matrices).
//   * product of digit planes p, q has weight 2^(-12-8(p+q)); all pairs with p+q = t <= 7 accumulate EXACTLY (int32,
//     |acc| <= 8 K 2^14: K <= 4096 per task) into TMEM accumulator t.  The dropped pairs (p+q >= 8) are the whole error:
//     ~2^-50 sqrt(K/4096) of sqrt(K_ii K_jj) -- the size of an FP64 GEMM's rounding error on the same data, far inside the
//     classical Cholesky backward-error bound (tests/test_ozaki_math.py is the NumPy specification).
//   * epilogue: the 8 accumulators of an output entry are combined in int64, converted once, scaled by 2^(E_i+E_j-68)
//     and subtracted from the target panel with RED.ADD.F64.
//
// One CTA = one 128 x 64 output tile (8 accumulators x 64 columns = all 512 TMEM columns), K streamed in 32-byte
// chunks (one MMA K-step of all 8 planes of both operands = 48 KiB per stage) by bulk async copies through a 4-stage
// mbarrier ring; warp 0 = copy producer, warp 1 = MMA issuer (one thread), warps 2-5 = epilogue (one TMEM lane quarter each).
// Because the planes of a row use ONE exponent for the whole row of L, any K range accumulates exactly: tasks are
// left-looking over many column pieces, so the RED epilogue is paid once per (tile, K range) instead of once per piece.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/tlpb200.h"
#include "kernels.cuh"

namespace tlp {

namespace {

constexpr int OZ_KB = 32;                    // K bytes (= int8 elements) per stage and per MMA
constexpr int OZ_BLK = 128 * OZ_KB;          // 4096 B: one plane of one 128-row block for one K chunk
constexpr int OZ_TN = 64;                    // output tile columns
constexpr int OZ_STAGE_A = OZ_S * OZ_BLK;            // 32 KiB
constexpr int OZ_STAGE_B = OZ_S * (OZ_BLK / 2);      // 16 KiB (64 rows)
constexpr int OZ_STAGE = OZ_STAGE_A + OZ_STAGE_B;    // 48 KiB
constexpr int OZ_NSTAGE = 4;
constexpr int OZ_THREADS = 192;
constexpr int OZ_TMEM_COLS = 512;
constexpr size_t OZ_SMEM = (size_t)OZ_NSTAGE * OZ_STAGE + 1024;   // + alignment slack

// core-matrix strides of the operand planes (bytes): 8 rows x 16 bytes contiguous; next 16 K-bytes at +128 (LBO);
// next 8 rows at +256 (SBO)
constexpr uint32_t OZ_LBO = 128, OZ_SBO = 256;

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) bits [4,6); a/b format INT8 (1) bits [7,10),
// [10,13); both K-major; N >> 3 bits [17,23); M >> 4 bits [24,29)
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(OZ_LBO >> 4) << 16) | ((uint64_t)(OZ_SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a broken pipeline raises err[0] instead of hanging the device
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int32_t* err) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t idesc = OZ_IDESC) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

struct OzShared {
    unsigned long long full[OZ_NSTAGE];
    unsigned long long empty[OZ_NSTAGE];
    unsigned long long acc_full;
    unsigned long long acc_empty;
    uint32_t tmem_base;
    int32_t next;
};

}  // namespace

// ------------------------------------------------------------------------------------------
// E_r of every panel row of one oz supernode, taken from the CURRENT diagonal entry of that row's column when the
// supernode starts (before its first piece is factored): d_r = K_rr - sum_{k done} L_rk^2 >= sum_{k in supernode} L_rk^2,
// so |L_rk| <= sqrt(d_r) < 2^E_r for every column k of the supernode -- a much tighter scale than sqrt(K_rr) deep in
// the elimination tree.  Any snapshot taken before the supernode's own updates land is a valid bound.
// ------------------------------------------------------------------------------------------
__global__ void k_oz_rowexp(const double* __restrict__ Lx, const int64_t* __restrict__ diagpos, const int32_t* __restrict__ grow,
                            int32_t nrows, int32_t* __restrict__ E, double* __restrict__ scl) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const double d = *reinterpret_cast<const volatile double*>(Lx + diagpos[grow[r]]);
    int e = 0;
    if (d > 0.0 && d < 1.7e308) frexp(d, &e); else e = 0;      // bad diagonals are caught by the factorisation itself
    const int Er = (e + 1) >> 1;
    E[r] = Er;
    scl[r] = ldexp(1.0, Er);
}

// ------------------------------------------------------------------------------------------
// digit planes of the columns [c0, c0 + w) (w <= 128, one piece) of a panel, rows [rb0*128, nrows) :
// grid (row blocks, 8 chunks of 16 columns), 128 threads = rows of the block
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_oz_slice(const double* __restrict__ panel, int64_t ld, int32_t nrows, int32_t c0, int32_t w,
                                                  int32_t rb0, int32_t kchunk0, const int32_t* __restrict__ E,
                                                  const int64_t* __restrict__ rb_off, uint8_t* __restrict__ planes) {
    const int32_t rb = rb0 + blockIdx.x;
    const int kc = blockIdx.y;                      // 16-column chunk inside the piece
    const int r = threadIdx.x;
    const int32_t row = rb * 128 + r;
    uint32_t pk[OZ_S][4];
#pragma unroll
    for (int p = 0; p < OZ_S; ++p) pk[p][0] = pk[p][1] = pk[p][2] = pk[p][3] = 0u;
    if (row < nrows) {
        const double sc = ldexp(1.0, 62 - E[row]);
        const double* src = panel + (int64_t)(c0 + kc * 16) * ld + row;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // |x 2^-E| < 1: 62-bit fixed point, then balanced base-256 digits (exact: q = sum_p d_p 256^(7-p))
            long long q = (kc * 16 + k < w) ? __double2ll_rn(src[(int64_t)k * ld] * sc) : 0ll;
#pragma unroll
            for (int p = OZ_S - 1; p >= 1; --p) {
                const int d = (int)(signed char)(q & 0xff);
                q = (q - d) >> 8;
                pk[p][k >> 2] |= ((uint32_t)(d & 0xff)) << (8 * (k & 3));
            }
            pk[0][k >> 2] |= ((uint32_t)((int)q & 0xff)) << (8 * (k & 3));
        }
    }
    uint8_t* dst = planes + (rb_off[rb] + (int64_t)(kchunk0 + (kc >> 1))) * (int64_t)(OZ_S * OZ_BLK) + (r >> 3) * OZ_SBO + (kc & 1) * OZ_LBO +
                   (r & 7) * 16;
#pragma unroll
    for (int p = 0; p < OZ_S; ++p) *reinterpret_cast<uint4*>(dst + p * OZ_BLK) = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
}

// ------------------------------------------------------------------------------------------
// tile update
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OZ_THREADS, 1) k_oz_update(const OzView* __restrict__ views, const OzTask* __restrict__ tasks, int32_t begin,
                                                             int32_t end, int32_t* counter, int32_t* err) {
    extern __shared__ uint8_t oz_smem_raw[];
    __shared__ OzShared sh;
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < OZ_NSTAGE; ++s) {
            mbar_init(smem_u32(&sh.full[s]), 1);
            mbar_init(smem_u32(&sh.empty[s]), 1);
        }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), 4);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sh.tmem_base)), "n"(OZ_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = sh.tmem_base;

    uint32_t it = 0;          // stages consumed / produced so far (same sequence in the producer and the MMA thread)
    uint32_t ntask = 0;       // tasks processed so far by this CTA
    bool alive = true;
    for (;;) {
        if (tid == 0) {
            const int32_t nx = begin + atomicAdd(counter, 1);
            sh.next = (*reinterpret_cast<volatile int32_t*>(err) != 0) ? end : nx;     // a timed-out pipeline drains the queue at once
        }
        __syncthreads();
        const int32_t task = sh.next;
        if (task >= end) break;
        const OzTask T = tasks[task];
        const OzView V = views[T.view];
        const int32_t nk = T.k1 - T.k0;

        if (tid == 0) {
            // ===== producer: all 8 planes of both operands for one K chunk per stage =====
            const uint8_t* a = V.planes + (V.rb_off[T.rbA] + (int64_t)T.k0) * (int64_t)OZ_STAGE_A;
            const uint8_t* b = V.planes + (V.rb_off[T.rbB] + (int64_t)T.k0) * (int64_t)OZ_STAGE_A + T.half * (OZ_BLK / 2);
            uint32_t i = it;
            for (int32_t kc = 0; kc < nk && alive; ++kc, ++i) {
                const uint32_t s = i % OZ_NSTAGE, ph = (i / OZ_NSTAGE) & 1u;
                alive = mbar_wait(smem_u32(&sh.empty[s]), ph ^ 1u, err);
                const uint32_t fb = smem_u32(&sh.full[s]);
                const uint32_t dst = stage0 + s * OZ_STAGE;
                mbar_expect_tx(fb, OZ_STAGE);
                bulk_g2s(dst, a + (int64_t)kc * OZ_STAGE_A, OZ_STAGE_A, fb);
#pragma unroll
                for (int p = 0; p < OZ_S; ++p)
                    bulk_g2s(dst + OZ_STAGE_A + p * (OZ_BLK / 2), b + (int64_t)kc * OZ_STAGE_A + p * OZ_BLK, OZ_BLK / 2, fb);
            }
        } else if (tid == 32) {
            // ===== MMA issuer =====
            alive = mbar_wait(smem_u32(&sh.acc_empty), (ntask & 1u) ^ 1u, err);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            uint32_t i = it;
            for (int32_t kc = 0; kc < nk && alive; ++kc, ++i) {
                const uint32_t s = i % OZ_NSTAGE, ph = (i / OZ_NSTAGE) & 1u;
                alive = mbar_wait(smem_u32(&sh.full[s]), ph, err);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t sa = stage0 + s * OZ_STAGE, sb = sa + OZ_STAGE_A;
#pragma unroll
                for (int p = 0; p < OZ_S; ++p) {
                    const uint64_t da = oz_desc(sa + p * OZ_BLK);
#pragma unroll
                    for (int q = 0; q + p < OZ_S; ++q) {
                        const uint64_t db = oz_desc(sb + q * (OZ_BLK / 2));
                        umma_i8(tmem + (uint32_t)((p + q) * OZ_TN), da, db, (kc > 0 || p > 0) ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&sh.empty[s]));      // frees the stage when these MMAs have read it
            }
            umma_commit(smem_u32(&sh.acc_full));
        } else if (warp >= 2) {
            // ===== epilogue: TMEM lane quarter (warp % 4), thread = output row =====
            alive = mbar_wait(smem_u32(&sh.acc_full), ntask & 1u, err);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int q4 = warp & 3;
            const int32_t i_loc = q4 * 32 + lane;
            const int32_t gi = T.rbA * 128 + i_loc;               // row inside the view
            const int32_t gj0 = T.rbB * 128 + T.half * OZ_TN;     // first column of the tile
            const bool rv = gi < V.nrows;
            const double si = rv ? V.scl[gi] * 0x1p-68 : 0.0;
            double* crow = V.C + gi;
            const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
            for (int cc = 0; cc < OZ_TN; cc += 8) {
                int32_t a[OZ_S][8];
#pragma unroll
                for (int t = 0; t < OZ_S; ++t) tmem_ld8(tbase + (uint32_t)(t * OZ_TN + cc), a[t]);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int32_t gj = gj0 + cc + e;
                    long long hi = a[0][e], lo = a[4][e];
#pragma unroll
                    for (int t = 1; t < 4; ++t) {
                        hi = hi * 256 + a[t][e];
                        lo = lo * 256 + a[4 + t][e];
                    }
                    if (rv && gj < V.ncols && gi >= gj) {
                        const double val = ((double)hi * 0x1p32 + (double)lo) * si * V.scl[gj];
                        atomicAdd(crow + (int64_t)gj * V.ldc, -val);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sh.acc_empty));
        }
        it += (uint32_t)nk;
        ntask++;
        __syncthreads();
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------------------------------
// 128 x 128 tile, two passes over K (TLPB200_OZAKI_TILE=128).  The 128 x 64 kernel above re-reads 6 KiB of shared memory
// per 32-cycle MMA (192 B/clk against the 128 B/clk the SM delivers), so the tensor pipe idles a third of the time.  With
// N = 128 an MMA reads 8 KiB per 64 cycles (128 B/clk); the 8 accumulators of a tile then need 1024 TMEM columns, so the
// levels are done in two passes that each own all 512 columns: pass 0 = levels 0..3 (10 plane pairs, planes 0..3 only),
// pass 1 = levels 4..7 (26 pairs).  Each pass drains into registers (16 epilogue warps x 32 columns), hands TMEM back and
// issues its REDs underneath the next pass's MMAs.
// ------------------------------------------------------------------------------------------
constexpr int OZ2_THREADS = 576;                  // producer warp, MMA warp, 16 epilogue warps (4 per TMEM lane quarter)
constexpr int OZ2_STAGE = 2 * OZ_STAGE_A;          // 64 KiB: A planes at 0, B planes at 32 KiB
constexpr int OZ2_NSTAGE = 3;
constexpr size_t OZ2_SMEM = (size_t)OZ2_NSTAGE * OZ2_STAGE + 1024;
constexpr uint32_t OZ2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(OZ2_THREADS, 1) k_oz_update2(const OzView* __restrict__ views, const OzTask* __restrict__ tasks, int32_t begin,
                                                               int32_t end, int32_t* counter, int32_t* err) {
    extern __shared__ uint8_t oz_smem_raw[];
    __shared__ OzShared sh;
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t stage0 = (raw + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < OZ2_NSTAGE; ++s) {
            mbar_init(smem_u32(&sh.full[s]), 1);
            mbar_init(smem_u32(&sh.empty[s]), 1);
        }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), 16);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sh.tmem_base)), "n"(OZ_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = sh.tmem_base;

    uint32_t it = 0;          // stages produced / consumed so far
    uint32_t npass = 0;       // passes done so far by this CTA (2 per task)
    bool alive = true;
    for (;;) {
        if (tid == 0) {
            const int32_t nx = begin + atomicAdd(counter, 1);
            sh.next = (*reinterpret_cast<volatile int32_t*>(err) != 0) ? end : nx;
        }
        __syncthreads();
        const int32_t task = sh.next;
        if (task >= end) break;
        const OzTask T = tasks[task];
        const OzView V = views[T.view];
        const int32_t nk = T.k1 - T.k0;

        if (tid == 0) {
            // ===== producer: pass 0 streams planes 0..3 of both operands, pass 1 all 8 =====
            const uint8_t* a = V.planes + (V.rb_off[T.rbA] + (int64_t)T.k0) * (int64_t)OZ_STAGE_A;
            const uint8_t* b = V.planes + (V.rb_off[T.rbB] + (int64_t)T.k0) * (int64_t)OZ_STAGE_A;
            uint32_t i = it;
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t bytes = pass == 0 ? OZ_STAGE_A / 2 : OZ_STAGE_A;
                for (int32_t kc = 0; kc < nk && alive; ++kc, ++i) {
                    const uint32_t s = i % OZ2_NSTAGE, ph = (i / OZ2_NSTAGE) & 1u;
                    alive = mbar_wait(smem_u32(&sh.empty[s]), ph ^ 1u, err);
                    const uint32_t fb = smem_u32(&sh.full[s]);
                    const uint32_t dst = stage0 + s * OZ2_STAGE;
                    mbar_expect_tx(fb, 2 * bytes);
                    bulk_g2s(dst, a + (int64_t)kc * OZ_STAGE_A, bytes, fb);
                    bulk_g2s(dst + OZ_STAGE_A, b + (int64_t)kc * OZ_STAGE_A, bytes, fb);
                }
            }
        } else if (tid == 32) {
            // ===== MMA issuer =====
            uint32_t i = it;
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t np = npass + pass;
                alive = alive && mbar_wait(smem_u32(&sh.acc_empty), (np & 1u) ^ 1u, err);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                for (int32_t kc = 0; kc < nk && alive; ++kc, ++i) {
                    const uint32_t s = i % OZ2_NSTAGE, ph = (i / OZ2_NSTAGE) & 1u;
                    alive = mbar_wait(smem_u32(&sh.full[s]), ph, err);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sa = stage0 + s * OZ2_STAGE, sb = sa + OZ_STAGE_A;
                    if (pass == 0) {
#pragma unroll
                        for (int p = 0; p < 4; ++p)
#pragma unroll
                            for (int q = 0; q + p < 4; ++q)
                                umma_i8(tmem + (uint32_t)((p + q) * 128), oz_desc(sa + p * OZ_BLK), oz_desc(sb + q * OZ_BLK),
                                        (kc > 0 || p > 0) ? 1u : 0u, OZ2_IDESC);
                    } else {
#pragma unroll
                        for (int p = 0; p < OZ_S; ++p)
#pragma unroll
                            for (int q = 0; q + p < OZ_S; ++q)
                                if (p + q >= 4)
                                    umma_i8(tmem + (uint32_t)((p + q - 4) * 128), oz_desc(sa + p * OZ_BLK), oz_desc(sb + q * OZ_BLK),
                                            (kc > 0 || p > 0) ? 1u : 0u, OZ2_IDESC);
                    }
                    umma_commit(smem_u32(&sh.empty[s]));
                }
                umma_commit(smem_u32(&sh.acc_full));
            }
        } else if (warp >= 2) {
            // ===== epilogue: 16 warps; TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4; thread = output row =====
            const int q4 = warp & 3, ch = (warp - 2) >> 2;
            const int32_t gi = T.rbA * 128 + q4 * 32 + lane;
            const int32_t gj0 = T.rbB * 128 + ch * 32;
            const bool rv = gi < V.nrows;
            const double sr = rv ? V.scl[gi] : 0.0;
            double* crow = V.C + gi;
            const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * 32);
            for (int pass = 0; pass < 2; ++pass) {
                alive = alive && mbar_wait(smem_u32(&sh.acc_full), (npass + pass) & 1u, err);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                double val[32];
#pragma unroll
                for (int cc = 0; cc < 32; cc += 8) {
                    int32_t a[4][8];
#pragma unroll
                    for (int t = 0; t < 4; ++t) tmem_ld8(tbase + (uint32_t)(t * 128 + cc), a[t]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        long long v = a[0][e];
#pragma unroll
                        for (int t = 1; t < 4; ++t) v = v * 256 + a[t][e];
                        val[cc + e] = (double)v;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sh.acc_empty));      // TMEM is free: the next pass starts underneath the REDs
                if (rv) {
                    const double si = sr * (pass == 0 ? 0x1p-36 : 0x1p-68);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int32_t gj = gj0 + e;
                        if (gj < V.ncols && gi >= gj) atomicAdd(crow + (int64_t)gj * V.ldc, -(val[e] * si * V.scl[gj]));
                    }
                }
            }
        }
        it += 2u * (uint32_t)nk;
        npass += 2u;
        __syncthreads();
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(OZ_TMEM_COLS) : "memory");
    }
}

cudaError_t ozaki_static_init() {
    cudaError_t e = cudaFuncSetAttribute(k_oz_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_oz_update2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OZ2_SMEM);
}

void launch_oz_rowexp(const double* Lx, const int64_t* diagpos, const int32_t* grow, int32_t nrows, int32_t* E, double* scl,
                      cudaStream_t st) {
    if (nrows > 0) k_oz_rowexp<<<(nrows + 255) / 256, 256, 0, st>>>(Lx, diagpos, grow, nrows, E, scl);
}

void launch_oz_slice(const double* panel, int64_t ld, int32_t nrows, int32_t c0, int32_t w, int32_t rb0, int32_t nrb, int32_t kchunk0,
                     const int32_t* E, const int64_t* rb_off, uint8_t* planes, cudaStream_t st) {
    if (nrb > 0) k_oz_slice<<<dim3((unsigned)nrb, 8), 128, 0, st>>>(panel, ld, nrows, c0, w, rb0, kchunk0, E, rb_off, planes);
}

void launch_oz_update(const OzView* views, const OzTask* tasks, int32_t begin, int32_t end, int32_t* counter, int nsm, int reserve,
                      int32_t* err, int tile_n, cudaStream_t st) {
    if (end <= begin) return;
    // One CTA owns a whole SM (192 KiB of shared memory, all of TMEM).  SMs are kept free for the critical-chain kernels by
    // the grid size, not by SM id: a CTA that exits because of where it landed is replaced by the next pending CTA of the
    // same grid on the same SM, and when the other SMs are busy the whole grid drains through the reserved ones.
    const int grid = std::min(end - begin, std::max(1, nsm - reserve));
    if (tile_n == 128) k_oz_update2<<<grid, OZ2_THREADS, OZ2_SMEM, st>>>(views, tasks, begin, end, counter, err);
    else k_oz_update<<<grid, OZ_THREADS, OZ_SMEM, st>>>(views, tasks, begin, end, counter, err);
}

}  // namespace tlp

// ------------------------------------------------------------------------------------------
// stand-alone check / timing of the path on a dense matrix (tests/test_gpu_ozaki.py, scripts/ozaki_probe.py)
// ------------------------------------------------------------------------------------------
extern "C" int tlpb200_debug_ozaki(const double* P, int64_t R, int64_t K, double* C, int32_t ksplit, int32_t reps, float* ms,
                                   int32_t* err_out) {
    using namespace tlp;
    if (!P || !C || R <= 0 || K <= 0 || !ms || !err_out) return TLPB200_BAD_ARG;
    int dev = 0, cc_major = 0, nsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return TLPB200_CUDA;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (cc_major != 10) return TLPB200_CUDA;
    if (ozaki_static_init() != cudaSuccess) return TLPB200_CUDA;
    const int32_t nrb = (int32_t)((R + 127) / 128);
    const int32_t npiece = (int32_t)((K + 127) / 128);
    const int32_t nk32 = npiece * 4;
    if (ksplit <= 0 || ksplit > 4096) ksplit = std::min(nk32 * 32, 4096);   // exact int32 accumulation needs K <= 4096 per task
    const int32_t kstep = std::max(1, ksplit / 32);
    std::vector<int64_t> rb_off(nrb + 1);
    for (int32_t r = 0; r <= nrb; ++r) rb_off[r] = (int64_t)r * nk32;
    std::vector<int32_t> E(R);
    std::vector<double> scl(R);
    for (int64_t r = 0; r < R; ++r) {          // as in production: from the diagonal of K = P P' (sqrt(K_rr) < 2^E)
        double d = 0.0;
        for (int64_t k = 0; k < K; ++k) d += P[k * R + r] * P[k * R + r];
        int e = 0;
        if (d > 0.0) frexp(d, &e);
        E[r] = (e + 1) >> 1;
        scl[r] = ldexp(1.0, E[r]);
    }
    const int tile_n = (getenv("TLPB200_OZAKI_TILE") && atoi(getenv("TLPB200_OZAKI_TILE")) == 128) ? 128 : 64;
    std::vector<OzTask> tasks;
    for (int32_t b = 0; b < nrb; ++b)
        for (int32_t a = b; a < nrb; ++a)
            for (int32_t h = 0; h < (tile_n == 128 ? 1 : 2); ++h)
                for (int32_t k0 = 0; k0 < nk32; k0 += kstep) {
                    OzTask t{};
                    t.view = 0; t.rbA = a; t.rbB = b; t.half = h; t.k0 = k0; t.k1 = std::min(nk32, k0 + kstep);
                    tasks.push_back(t);
                }
    double *dP = nullptr, *dC = nullptr, *dscl = nullptr;
    uint8_t* dpl = nullptr;
    int64_t* doff = nullptr;
    int32_t *dE = nullptr, *dctr = nullptr, *derr = nullptr;
    OzView* dview = nullptr;
    OzTask* dtask = nullptr;
    const size_t plane_bytes = (size_t)nrb * nk32 * OZ_S * 4096;
    int rc = TLPB200_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
#define OZCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "tlpb200_debug_ozaki: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); rc = TLPB200_CUDA; goto done; } } while (0)
    OZCK(cudaMalloc(&dP, (size_t)R * K * 8));
    OZCK(cudaMalloc(&dC, (size_t)R * R * 8));
    OZCK(cudaMalloc(&dscl, (size_t)R * 8));
    OZCK(cudaMalloc(&dpl, plane_bytes));
    OZCK(cudaMalloc(&doff, (size_t)(nrb + 1) * 8));
    OZCK(cudaMalloc(&dE, (size_t)R * 4));
    OZCK(cudaMalloc(&dctr, 8));
    OZCK(cudaMalloc(&derr, 4));
    OZCK(cudaMalloc(&dview, sizeof(OzView)));
    OZCK(cudaMalloc(&dtask, tasks.size() * sizeof(OzTask)));
    OZCK(cudaMemcpy(dP, P, (size_t)R * K * 8, cudaMemcpyHostToDevice));
    OZCK(cudaMemcpy(dC, C, (size_t)R * R * 8, cudaMemcpyHostToDevice));
    OZCK(cudaMemcpy(dscl, scl.data(), (size_t)R * 8, cudaMemcpyHostToDevice));
    OZCK(cudaMemcpy(dE, E.data(), (size_t)R * 4, cudaMemcpyHostToDevice));
    OZCK(cudaMemcpy(doff, rb_off.data(), (size_t)(nrb + 1) * 8, cudaMemcpyHostToDevice));
    OZCK(cudaMemcpy(dtask, tasks.data(), tasks.size() * sizeof(OzTask), cudaMemcpyHostToDevice));
    OZCK(cudaMemset(derr, 0, 4));
    {
        OzView v{};
        v.planes = dpl; v.rb_off = doff; v.C = dC; v.ldc = R; v.scl = dscl; v.nrows = (int32_t)R; v.ncols = (int32_t)R;
        OZCK(cudaMemcpy(dview, &v, sizeof v, cudaMemcpyHostToDevice));
    }
    OZCK(cudaEventCreate(&e0));
    OZCK(cudaEventCreate(&e1));
    OZCK(cudaEventRecord(e0));
    for (int32_t j = 0; j < npiece; ++j)
        launch_oz_slice(dP, R, (int32_t)R, j * 128, (int32_t)std::min<int64_t>(128, K - (int64_t)j * 128), 0, nrb, 4 * j, dE, doff, dpl, 0);
    OZCK(cudaEventRecord(e1));
    OZCK(cudaMemset(dctr, 0, 8));
    launch_oz_update(dview, dtask, 0, (int32_t)tasks.size(), dctr, nsm, 0, derr, tile_n, 0);
    OZCK(cudaDeviceSynchronize());
    OZCK(cudaEventElapsedTime(&ms[0], e0, e1));
    OZCK(cudaMemcpy(C, dC, (size_t)R * R * 8, cudaMemcpyDeviceToHost));
    OZCK(cudaMemcpy(err_out, derr, 4, cudaMemcpyDeviceToHost));
    ms[1] = 0.f;
    if (reps > 0 && *err_out == 0) {
        OZCK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r) {
            OZCK(cudaMemsetAsync(dctr, 0, 8, 0));
            launch_oz_update(dview, dtask, 0, (int32_t)tasks.size(), dctr, nsm, 0, derr, tile_n, 0);
        }
        OZCK(cudaEventRecord(e1));
        OZCK(cudaDeviceSynchronize());
        OZCK(cudaEventElapsedTime(&ms[1], e0, e1));
        ms[1] /= (float)reps;
        OZCK(cudaMemcpy(err_out, derr, 4, cudaMemcpyDeviceToHost));
    }
    ms[2] = (float)tasks.size();
done:
    cudaFree(dP); cudaFree(dC); cudaFree(dscl); cudaFree(dpl); cudaFree(doff); cudaFree(dE); cudaFree(dctr); cudaFree(derr);
    cudaFree(dview); cudaFree(dtask);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
#undef OZCK
}
