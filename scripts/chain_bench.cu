// Phase timeline of the critical-chain kernels on one synthetic 128-column piece (round 2, session Q):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTLP_CHAIN_CLOCKS -I tulip.jl_b200/csrc scripts/chain_bench.cu -o scripts/chain_bench
// Prints, per kernel, the CUDA-event time of a launch and the clock64 deltas between the TLP_TICK marks of CTA 0.
#include "../tulip.jl_b200/csrc/kernels_factor.cu"

#include <cmath>
#include <cstdio>
#include <vector>

using namespace tlp;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <class T>
T* up(const std::vector<T>& v) {
    T* p;
    cudaMalloc(&p, v.size() * sizeof(T));
    cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}

static void ticks(const char* name, int n0, int n1) {
    unsigned long long h[2][80];
    cudaMemcpyFromSymbol(h, g_ticks, sizeof(h));
    printf("%s: clock64 total (thread 0) %llu cyc\n  marks:", name, h[0][n1] - h[0][n0]);
    for (int i = n0 + 1; i <= n1; ++i)
        if (h[0][i] > h[0][n0] && h[0][i] - h[0][n0] < 10000000ull) printf(" [%d] %llu/%llu", i, h[0][i] - h[0][n0], h[1][i] > h[1][n0] ? h[1][i] - h[1][n0] : 0ull);
    printf("\n");
}

__global__ void k_rcp_test(const double* x, double* worst, int n) {
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = x[i], r = fast_rcp(d), q = 1.0 / d;
        m = fmax(m, fabs(r - q) / fabs(q));
    }
    atomicMax((unsigned long long*)worst, (unsigned long long)__double_as_longlong(m));
}

// dependent-chain latency / single-warp throughput of the FP64 operations the chain kernels are made of
__global__ void k_lat(double* out, unsigned long long* cyc, double seed) {
    const int lane = threadIdx.x & 31;
    double a = seed + lane * 1e-3, b = 1.0 + seed * 1e-9, c = seed * 1e-7;
    unsigned long long t0, t1;
    // 0: 256 dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 1: 8 independent chains x 32
    double v[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) v[x] = a + x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int x = 0; x < 8; ++x) v[x] = fma(v[x], b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
#pragma unroll
    for (int x = 0; x < 8; ++x) a += v[x];
    // 2: 64 dependent rcp.approx.ftz.f64 (+ one DFMA each to keep the value in range)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
        a = fma(r, b, c);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 3: 64 dependent DMMA
    double c0 = a, c1 = a * 0.5;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) dmma884(c0, c1, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 4: 8 independent DMMA chains x 8
    double w0[8], w1[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) { w0[x] = c0 + x; w1[x] = c1 - x; }
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int x = 0; x < 8; ++x) dmma884(w0[x], w1[x], b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
#pragma unroll
    for (int x = 0; x < 8; ++x) a += w0[x] + w1[x];
    // 5: 64 dependent full-precision divisions 1.0 / a
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) a = 1.0 / (a + 1.5);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 6: 64 dependent shuffles
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 7: 64 dependent shared-memory round trips (store + load)
    __shared__ double sh[64];
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        sh[lane] = a;
        __syncwarp();
        a = sh[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    // 8: 256 dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = a * b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[8] = t1 - t0;
    out[threadIdx.x] = a + c0 + c1;
}

// FP64 issue rate: every warp runs 8 independent chains of 64 DFMAs (or DMMAs); nw warps, i.e. nw / 4 per SMSP
__global__ void k_tput(double* out, unsigned long long* cyc, double seed, int mode) {
    double v[8], w[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) { v[x] = seed + x + threadIdx.x * 1e-3; w[x] = seed - x; }
    const double b = 1.0 + seed * 1e-9, c = seed * 1e-7;
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (mode == 0) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
#pragma unroll
            for (int x = 0; x < 8; ++x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[x]) : "d"(b), "d"(c));
    } else {
#pragma unroll
        for (int i = 0; i < 64; ++i)
#pragma unroll
            for (int x = 0; x < 8; ++x) dmma884(v[x], w[x], b, c);
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double a = 0;
#pragma unroll
    for (int x = 0; x < 8; ++x) a += v[x] + w[x];
    out[threadIdx.x] = a;
}

int main() {
    {
        double* o; unsigned long long* cy;
        cudaMalloc(&o, 1024 * 8); cudaMalloc(&cy, 16 * 8);
        for (int mode = 0; mode < 2; ++mode)
            for (int nw = 1; nw <= 16; nw *= 2) {
                k_tput<<<1, 32 * nw>>>(o, cy, 1.000001, mode);
                cudaDeviceSynchronize();
                k_tput<<<1, 32 * nw>>>(o, cy, 1.000001, mode);
                unsigned long long h = 0;
                cudaMemcpy(&h, cy, 8, cudaMemcpyDeviceToHost);
                printf("k_tput %s, %2d warps in one CTA: %.2f cycles per instruction per warp (512 independent-chain instructions, warp 0's clock)\n",
                       mode ? "DMMA" : "DFMA", nw, h / 512.0);
            }
    }
    {
        double* o; unsigned long long* cy;
        cudaMalloc(&o, 1024 * 8); cudaMalloc(&cy, 16 * 8);
        for (int nw = 1; nw <= 4; nw *= 4) {
            k_lat<<<1, 32 * nw>>>(o, cy, 1.000001);
            cudaDeviceSynchronize();
            k_lat<<<1, 32 * nw>>>(o, cy, 1.000001);
            unsigned long long h[16];
            cudaMemcpy(h, cy, sizeof(h), cudaMemcpyDeviceToHost);
            printf("k_lat, %d warp(s) (one per SMSP): DFMA dependent %.1f cyc, 8 chains %.2f cyc/instr; rcp.approx+DFMA %.1f; DMMA dependent %.1f, 8 chains %.2f cyc/instr; "
                   "1.0/x (+DADD) %.1f; shfl %.1f; smem round trip (+DADD) %.1f; DMUL dependent %.1f\n",
                   nw, h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 64.0, h[8] / 256.0);
        }
    }
    {   // fast_rcp against the correctly rounded quotient
        const int n = 1 << 22;
        std::vector<double> x(n);
        unsigned long long sd = 99;
        for (int i = 0; i < n; ++i) {
            sd = sd * 6364136223846793005ull + 1442695040888963407ull;
            const double u = (double)(sd >> 11) / 9007199254740992.0;
            x[i] = (i & 1 ? -1.0 : 1.0) * std::ldexp(1.0 + u, (int)((sd >> 3) % 600) - 300);
        }
        double* dx = up(x);
        std::vector<double> z(1, 0.0);
        double* dw = up(z);
        k_rcp_test<<<296, 256>>>(dx, dw, n);
        cudaMemcpy(z.data(), dw, 8, cudaMemcpyDeviceToHost);
        printf("fast_rcp: max relative error vs 1.0/d over %d values = %.3e (2^-53 = 1.11e-16)\n", n, z[0]);
    }
    const int w = 128, nrow = 128 + 1024;
    std::vector<int32_t> sn_first = {0, 2 * w};
    std::vector<int64_t> sn_rowptr = {0, nrow}, sn_xptr = {0, (int64_t)nrow * 2 * w};
    std::vector<int32_t> sn_rows(nrow);
    for (int i = 0; i < nrow; ++i) sn_rows[i] = i;
    std::vector<int8_t> sign(nrow, 1);
    std::vector<Piece> pieces = {{0, 0, w, 0}, {0, w, 2 * w, 1}};
    std::vector<UpdTask> upd(3);
    upd[0].piece = 0; upd[0].i0 = 128; upd[0].ni = 64; upd[0].k0 = 128; upd[0].nk = 64; upd[0].tgt = 0; upd[0].diag = 1;
    upd[1] = upd[0]; upd[1].i0 = 192; upd[1].diag = 0;
    upd[2] = upd[0]; upd[2].i0 = 192; upd[2].k0 = 192;
    std::vector<int32_t> level_pieces = {0};
    std::vector<PanelTask> panel;
    for (int r = w; r < nrow; r += 128) panel.push_back({0, r, std::min(128, nrow - r), 0});
    // SPD diagonal block + random rows below
    std::vector<double> L0((size_t)nrow * 2 * w, 0.0), K((size_t)nrow * 2 * w, 0.0);
    unsigned long long sd = 12345;
    auto rnd = [&] { sd = sd * 6364136223846793005ull + 1442695040888963407ull; return (double)(sd >> 11) / 9007199254740992.0 - 0.5; };
    for (int k = 0; k < w; ++k)
        for (int i = k; i < nrow; ++i) L0[(size_t)k * nrow + i] = (i == k) ? 1.5 + rnd() : 0.4 * rnd();
    for (int k = 0; k < w; ++k)      // K = L0 L0' (columns 0..w-1, rows k..nrow-1)
        for (int i = k; i < nrow; ++i) {
            double a = 0;
            for (int j = 0; j <= k; ++j) a += L0[(size_t)j * nrow + i] * L0[(size_t)j * nrow + k];
            K[(size_t)k * nrow + i] = a;
        }
    DevCtx c{};
    c.sn_first = up(sn_first); c.sn_rowptr = up(sn_rowptr); c.sn_rows = up(sn_rows); c.sn_xptr = up(sn_xptr); c.sign = up(sign);
    c.pieces = up(pieces); c.level_pieces = up(level_pieces); c.panel = up(panel); c.upd = up(upd);
    double* dK = up(K);
    CK(cudaMalloc(&c.Lx, K.size() * 8));
    std::vector<int32_t> info(4, 0x7f7f7f7f);
    c.info = up(info);
    CK(factor_kernels_static_init());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<double> out(K.size());
    auto check = [&](const char* name) {
        cudaMemcpy(out.data(), c.Lx, out.size() * 8, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int k = 0; k < w; ++k)
            for (int i = k; i < nrow; ++i) err = std::max(err, std::fabs(out[(size_t)k * nrow + i] - L0[(size_t)k * nrow + i]));
        printf("%s: max |L - L0| = %.3e\n", name, err);
    };
    for (int variant = 0; variant < 2; ++variant) {
        float best_d = 1e9f, best_t = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaMemcpy(c.Lx, dK, K.size() * 8, cudaMemcpyDeviceToDevice));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            if (variant == 0) k_diag_factor<<<1, DF_THREADS, DF_SMEM>>>(c, 0);
            else k_diag_factor2<<<1, DF2_THREADS, DF2_SMEM>>>(c, 0);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); best_d = std::min(best_d, ms);
            if (rep == 5) { char nm[64]; snprintf(nm, 64, "diag variant %d", variant); ticks(nm, 0, 72); }
            cudaEventRecord(e0);
            if (variant == 0) k_trsm<<<(int)panel.size(), TR_THREADS, TR_SMEM>>>(c, 0);
            else k_trsm2<<<2 * (int)panel.size(), TR2_THREADS, TR2_SMEM>>>(c, 0);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, e0, e1); best_t = std::min(best_t, ms);
            if (rep == 5) { char nm[64]; snprintf(nm, 64, "trsm variant %d", variant); ticks(nm, 0, 20); }
            if (variant == 0) {
                for (int ks = 1; ks <= 8; ks *= 2) {
                    cudaEventRecord(e0);
                    k_update<<<3 * ks, UPD_THREADS, UPD_SMEM>>>(c, 0, 1, ks);
                    cudaEventRecord(e1);
                    CK(cudaDeviceSynchronize());
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep == 5) { printf("k_update, 3 critical 64x64x128 tiles, K split over %d CTAs: %.2f us (CUDA events)\n", ks, ms * 1e3f); ticks("k_update CTA 0", 0, 2); }
                }
            }
        }
        printf("variant %d: diag %.2f us, trsm %.2f us (CUDA events, best of 6)\n", variant, best_d * 1e3f, best_t * 1e3f);
        char nm[64]; snprintf(nm, 64, "variant %d", variant); check(nm);
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock attribute %d kHz\n", clk);
    return 0;
}
