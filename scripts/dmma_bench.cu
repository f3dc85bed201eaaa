// Microbenchmark (dev tool): FP64 throughput of DFMA and of the mma.sync f64 shapes on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void bench(double* out, int iters) {
    double c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = 0.5 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (SHAPE == 0) {        // plain DFMA: 8 independent chains x 4
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fma(a[i], b[j], c[i][j]);
            } else if (SHAPE == 884) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[0]));
            } else if (SHAPE == 1684) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[i]), "d"(a[(i + 1) & 7]), "d"(b[0]));
            } else if (SHAPE == 1688) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[i]), "d"(a[(i + 1) & 7]), "d"(a[(i + 2) & 7]), "d"(a[(i + 3) & 7]), "d"(b[0]), "d"(b[1]));
            } else if (SHAPE == 16816) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double fma_per_instr_warp, int warps_per_sm) {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 32 * warps_per_sm, blocks = p.multiProcessorCount, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    bench<SHAPE><<<blocks, threads>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<SHAPE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)iters * 8 * (SHAPE == 0 ? 4 : 1) * warps_per_sm * blocks;   // warp-level instructions
    const double fma = instr * fma_per_instr_warp;
    printf("%-10s warps/SM=%2d  %.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz)\n", name, warps_per_sm, 2 * fma / (ms * 1e-3) / 1e12,
           fma / (ms * 1e-3) / blocks / (p.clockRate * 1e3), p.clockRate / 1000);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("DFMA", 32, w);
        run<884>("m8n8k4", 256, w);
        run<1684>("m16n8k4", 512, w);
        run<1688>("m16n8k8", 1024, w);
        run<16816>("m16n8k16", 2048, w);
    }
    return 0;
}
