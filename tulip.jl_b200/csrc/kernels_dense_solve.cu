// Dense-solve path of the big supernodes (sm_100a): the triangular sweeps of F \ xi
// (/root/reference/src/KKT/Cholmod/spd.jl:61, sqd.jl:66) over a supernode with many 128-column blocks
// -- the dense root of a filled-in normal-equations factor -- are a chain of dependent block steps.
// The chain is made as short as the hardware allows:
//
//  * after the factorisation the panel is repacked (k_pack_big) into contiguous 128x128 tiles of
//        Lhat = L * blockdiag(L_kk)^{-1}          (unit block diagonal)
//    so that  L u = b  <=>  Lhat w = b,  u = blockdiag(L_kk)^{-1} w : no diagonal-block solve is left on
//    the chain; the block-diagonal factors are applied off the chain at the start of each backward task
//    (z_k = L_kk^{-T} S L_kk^{-1} w_k, then Lhat' x = z).  Ft holds the tiles column-major (forward sweep,
//    thread = output row), Bt row-major (backward sweep, thread = output column): both sweeps stream their
//    tiles with the same fully coalesced access pattern and every tile of L is read once per sweep.
//  * one persistent CTA per block row (forward) / block column (backward) streams its tiles in lock-step
//    with the chain.  Solved blocks are handed over through "flag-in-data" exchange slots
//    {bits(x), bits(x) ^ key(sweep)}: one 16-byte store by the producer, one polling 16-byte load by every
//    consumer thread -- no fence, no separate flag, no reset between sweeps (the key changes every sweep).
//    A chain step is then: poll hit -> 32 DFMA per thread on a tile that is already in registers -> one
//    shared-memory reduction -> publish.
#include "kernels.cuh"

namespace tlp {

namespace {

constexpr int TILE_ELEMS = SBLK * SBLK;   // 16384 doubles = 128 KiB

__device__ __forceinline__ unsigned long long sweep_key(unsigned long long epoch) {
    return (epoch + 1ull) * 0x9E3779B97F4A7C15ull;   // odd multiplier: distinct epochs -> distinct non-zero keys
}

__device__ __forceinline__ void publish_slot(unsigned long long* slot, double v, unsigned long long key) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" ::"l"(slot), "l"(b), "l"(b ^ key) : "memory");
}

// spin until the slot carries this sweep's key; a torn or stale read fails the xor test and is retried.
// The spin count is bounded so that a broken hand-over can never hang the device (info[2] reports it).
__device__ __forceinline__ double poll_slot(const unsigned long long* slot, unsigned long long key, int32_t* info) {
    unsigned long long a, b;
    int spins = 0;
    while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(a), "=l"(b) : "l"(slot) : "memory");
        if ((a ^ b) == key) break;
        if (++spins > (1 << 22)) { atomicExch(info + 2, 1); break; }
    }
    return __longlong_as_double((long long)a);
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(p), "r"(bytes) : "memory");
}

// the last CTA to leave a sweep kernel bumps the sweep counter (every CTA has read it by then)
__device__ __forceinline__ void leave_sweep(unsigned long long* epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(epoch + 1, 1ull);
        if (done == (unsigned long long)gridDim.x - 1ull) {
            epoch[1] = 0ull;
            __threadfence();
            atomicAdd(epoch, 1ull);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// repack: one CTA per tile.  That = T * X with T = L[rows, block j], X = L_jj^{-1} (lower triangular,
// taken from DinvT: DinvT[r*128 + c] = X[r, c]).
// ------------------------------------------------------------------------------------------
constexpr int PK_THREADS = 256;
constexpr int PK_HALF = 64;
constexpr size_t PK_SMEM = ((size_t)PK_HALF * SBLK + (size_t)SBLK * SBLK) * 8;

__global__ void __launch_bounds__(PK_THREADS, 1) k_pack_big(DevCtx c, int32_t begin) {
    extern __shared__ double smem_pk[];
    double* Ts = smem_pk;                    // [r][i]  r = column inside block j, i = row inside the half
    double* Xs = Ts + PK_HALF * SBLK;        // [r][c]
    const BigPack P = c.big_pack[begin + blockIdx.x];
    const int32_t s = P.sn;
    if (c.skip && c.skip[s]) return;
    const int32_t f = c.sn_first[s];
    const int32_t nc = c.sn_first[s + 1] - f;
    const int32_t ld = (int32_t)(c.sn_rowptr[s + 1] - c.sn_rowptr[s]);
    const double* panel = c.Lx + c.sn_xptr[s];
    const int32_t nbj = min(SBLK, nc - P.j * SBLK);
    const int tid = threadIdx.x;
    {
        const double2* src = reinterpret_cast<const double2*>(c.DinvT + (int64_t)(c.sn_dblk[s] + P.j) * TILE_ELEMS);
        double2* dst = reinterpret_cast<double2*>(Xs);
        for (int e = tid; e < TILE_ELEMS / 2; e += PK_THREADS) dst[e] = src[e];
    }
    const int ti = tid & 15, tj = tid >> 4;
    double* fout = c.Ft + P.fdst * TILE_ELEMS;
    double* bout = c.Bt + P.bdst * TILE_ELEMS;
    for (int h = 0; h < 2; ++h) {
        __syncthreads();   // Xs ready (h = 0) / Ts free again (h = 1)
        {
            const int i = tid & (PK_HALF - 1), rs = tid >> 6;
            const int32_t row = h * PK_HALF + i;
            const bool rv = row < P.nr;
            const double* src = panel + (int64_t)(P.j * SBLK) * ld + P.r0 + (rv ? row : 0);
            for (int r0 = 0; r0 < SBLK; r0 += PK_THREADS / PK_HALF) {
                const int r = r0 + rs;
                Ts[r * PK_HALF + i] = (rv && r < nbj) ? src[(int64_t)r * ld] : 0.0;
            }
        }
        __syncthreads();
        double o[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) o[a][b] = 0.0;
        // X[r, c] = 0 for r < c: start at the first column of the warp's column pair (no divergence)
        for (int r = (tj & ~1) * 8; r < SBLK; ++r) {
            const double2 a01 = *reinterpret_cast<const double2*>(Ts + r * PK_HALF + ti * 4);
            const double2 a23 = *reinterpret_cast<const double2*>(Ts + r * PK_HALF + ti * 4 + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y};
            double xv[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 x = *reinterpret_cast<const double2*>(Xs + r * SBLK + tj * 8 + 2 * q);
                xv[2 * q] = x.x;
                xv[2 * q + 1] = x.y;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) o[a][b] += av[a] * xv[b];
        }
        const int row0 = h * PK_HALF + ti * 4;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double* d = fout + (tj * 8 + b) * SBLK + row0;
            *reinterpret_cast<double2*>(d) = make_double2(o[0][b], o[1][b]);
            *reinterpret_cast<double2*>(d + 2) = make_double2(o[2][b], o[3][b]);
        }
        if (P.bdst >= 0)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double* d = bout + (row0 + a) * SBLK + tj * 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<double2*>(d + 2 * q) = make_double2(o[a][2 * q], o[a][2 * q + 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// sweeps
// ------------------------------------------------------------------------------------------
constexpr int BG_THREADS = 512;
constexpr int BG_CG = BG_THREADS / SBLK;   // 4 column groups of 32

struct BgShared {
    double xs[2][SBLK];
    double red[BG_CG][SBLK];
};

// registers <- this thread's 32 entries of a tile: row r, columns cg*32 .. cg*32+31 (tile pattern [c*128 + r])
__device__ __forceinline__ void load_tile(double (&t)[32], const double* tile, int r, int cg) {
    const double* p = tile + (cg * 32) * SBLK + r;
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) t[cc] = __ldcs(p + cc * SBLK);
}

__device__ __forceinline__ double dot_tile(const double (&t)[32], const double* xs, int cg) {
    double a = 0.0;
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) a += t[cc] * xs[cg * 32 + cc];
    return a;
}

// forward:  w_i = b_i - sum_{j<i} Lhat_ij w_j   (kind 0) ;  b_below -= Lhat_below,j w_j  (kind 1)
__global__ void __launch_bounds__(BG_THREADS, 1) k_fwd_big(DevCtx c, int32_t begin, int32_t end) {
    __shared__ BgShared sh;
    const int tid = threadIdx.x, r = tid & (SBLK - 1), cg = tid >> 7;
    const unsigned long long key = sweep_key(*reinterpret_cast<volatile unsigned long long*>(c.epoch));
    for (int32_t it = begin + blockIdx.x; it < end; it += gridDim.x) {
        const BigTask T = c.fwd_big[it];
        const int32_t s = T.sn;
        if (c.skip && c.skip[s]) continue;
        const int32_t f = c.sn_first[s];
        const int64_t rp = c.sn_rowptr[s];
        const double* tiles = c.Ft + T.tile0 * TILE_ELEMS;
        unsigned long long* xq = c.xq + 2 * (int64_t)T.xq0;
        const double bown = (T.kind == 0 && tid < T.nr) ? __ldcg(c.wk + f + T.r0 + tid) : 0.0;
        double t[32];
        double acc = 0.0;
        if (T.ntile > 0) load_tile(t, tiles, r, cg);
        if (tid < 4 && T.ntile > 1) prefetch_l2(tiles + TILE_ELEMS + tid * (TILE_ELEMS / 4), TILE_ELEMS * 2);
        for (int32_t j = 0; j < T.ntile; ++j) {
            if (tid < 4 && j + 2 < T.ntile)
                prefetch_l2(tiles + (int64_t)(j + 2) * TILE_ELEMS + tid * (TILE_ELEMS / 4), TILE_ELEMS * 2);
            double* xs = sh.xs[j & 1];
            if (tid < SBLK) xs[tid] = poll_slot(xq + 2 * (j * SBLK + tid), key, c.info);
            __syncthreads();
            acc -= dot_tile(t, xs, cg);
            if (j + 1 < T.ntile) load_tile(t, tiles + (int64_t)(j + 1) * TILE_ELEMS, r, cg);
        }
        sh.red[cg][r] = acc;
        __syncthreads();
        if (tid < SBLK) {
            const double v = sh.red[0][tid] + sh.red[1][tid] + sh.red[2][tid] + sh.red[3][tid];
            if (T.kind == 0) {
                const double w = bown + v;   // rows beyond nr: zero tiles, bown = 0 -> publishes 0
                publish_slot(xq + 2 * (T.blk * SBLK + tid), w, key);
                if (tid < T.nr) c.wk[f + T.r0 + tid] = w;
                if (c.dbg_ts && tid == 0) c.dbg_ts[T.xq0 / SBLK + T.blk] = global_ns();
            } else if (tid < T.nr) {
                atomicAdd(c.wk + c.sn_rows[rp + T.r0 + tid], v);
            }
        }
        __syncthreads();   // sh is reused by the next task
    }
    leave_sweep(c.epoch);
}

// backward:  z_k = L_kk^{-T} (S L_kk^{-1} w_k - p_k) ;  x_k = z_k - sum_{j>k} Lhat_jk' x_j
// (p_k = contribution of the rows below the supernode's columns, accumulated by k_bwd_below on the unscaled panel)
__global__ void __launch_bounds__(BG_THREADS, 1) k_bwd_big(DevCtx c, int32_t begin, int32_t end) {
    __shared__ BgShared sh;
    const int tid = threadIdx.x, r = tid & (SBLK - 1), cg = tid >> 7;
    const unsigned long long key = sweep_key(*reinterpret_cast<volatile unsigned long long*>(c.epoch));
    for (int32_t it = begin + blockIdx.x; it < end; it += gridDim.x) {
        const BigTask T = c.bwd_big[it];
        const int32_t s = T.sn;
        if (c.skip && c.skip[s]) continue;
        const int32_t f = c.sn_first[s];
        const int32_t nc = c.sn_first[s + 1] - f;
        const int32_t ncb = (nc + SBLK - 1) / SBLK;
        const int32_t k = T.blk;
        const double* tiles = c.Bt + T.tile0 * TILE_ELEMS;
        unsigned long long* xq = c.xq + 2 * (int64_t)T.xq0;
        double t[32];
        // block-diagonal part, off the chain: only needs the forward result of this block
        const int64_t db = (int64_t)(c.sn_dblk[s] + k) * TILE_ELEMS;
        load_tile(t, c.Dinv + db, r, cg);
        if (tid < 8 && (tid >> 2) < T.ntile) prefetch_l2(tiles + tid * (TILE_ELEMS / 4), TILE_ELEMS * 2);   // tiles 0, 1
        if (tid < SBLK) sh.xs[0][tid] = (tid < T.nr) ? __ldcg(c.wk + f + k * SBLK + tid) : 0.0;
        __syncthreads();
        sh.red[cg][r] = dot_tile(t, sh.xs[0], cg);
        load_tile(t, c.DinvT + db, r, cg);
        __syncthreads();
        if (tid < SBLK) {   // y = S L_kk^{-1} w_k - (rows below the columns, accumulated by k_bwd_below; re-zeroed here)
            double y = 0.0;
            if (tid < T.nr) {
                y = (double)c.sign[f + k * SBLK + tid] * (sh.red[0][tid] + sh.red[1][tid] + sh.red[2][tid] + sh.red[3][tid]) -
                    __ldcg(c.bacc + f + k * SBLK + tid);
                c.bacc[f + k * SBLK + tid] = 0.0;
            }
            sh.xs[1][tid] = y;
        }
        __syncthreads();
        sh.red[cg][r] = dot_tile(t, sh.xs[1], cg);
        if (T.ntile > 0) load_tile(t, tiles, r, cg);
        __syncthreads();
        double zown = 0.0;
        if (tid < SBLK) zown = sh.red[0][tid] + sh.red[1][tid] + sh.red[2][tid] + sh.red[3][tid];
        __syncthreads();   // red / xs free again
        double acc = 0.0;
        for (int32_t j = 0; j < T.ntile; ++j) {
            if (tid < 4 && j + 2 < T.ntile)
                prefetch_l2(tiles + (int64_t)(j + 2) * TILE_ELEMS + tid * (TILE_ELEMS / 4), TILE_ELEMS * 2);
            double* xs = sh.xs[j & 1];
            if (tid < SBLK) xs[tid] = poll_slot(xq + 2 * ((ncb - 1 - j) * SBLK + tid), key, c.info);
            __syncthreads();
            acc -= dot_tile(t, xs, cg);
            if (j + 1 < T.ntile) load_tile(t, tiles + (int64_t)(j + 1) * TILE_ELEMS, r, cg);
        }
        sh.red[cg][r] = acc;
        __syncthreads();
        if (tid < SBLK) {
            const double x = zown + sh.red[0][tid] + sh.red[1][tid] + sh.red[2][tid] + sh.red[3][tid];
            publish_slot(xq + 2 * (k * SBLK + tid), x, key);
            if (tid < T.nr) c.wk[f + k * SBLK + tid] = x;
            if (c.dbg_ts && tid == 0) c.dbg_ts[c.nxblk + T.xq0 / SBLK + k] = global_ns();
        }
        __syncthreads();
    }
    leave_sweep(c.epoch);
}

// ------------------------------------------------------------------------------------------
cudaError_t dense_solve_static_init() {
    return cudaFuncSetAttribute(k_pack_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM);
}

void launch_pack_big(const DevCtx& c, int32_t begin, int32_t end, cudaStream_t st) {
    if (end > begin) k_pack_big<<<(unsigned)(end - begin), PK_THREADS, PK_SMEM, st>>>(c, begin);
}
void launch_fwd_big(const DevCtx& c, int32_t begin, int32_t end, int nsm, cudaStream_t st) {
    if (end > begin) k_fwd_big<<<(unsigned)min(end - begin, nsm), BG_THREADS, 0, st>>>(c, begin, end);
}
void launch_bwd_big(const DevCtx& c, int32_t begin, int32_t end, int nsm, cudaStream_t st) {
    if (end > begin) k_bwd_big<<<(unsigned)min(end - begin, nsm), BG_THREADS, 0, st>>>(c, begin, end);
}

}  // namespace tlp
