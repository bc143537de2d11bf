// Micro-benchmark: dependent-issue latency and per-SMSP throughput of the fp64 / select
// instruction classes k_stream4 is made of (B200, sm_100a).  nvcc -arch=sm_100a -O3 lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int n) {
    double x = a + threadIdx.x, y = b, z = a * 0.5, w = b * 2.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = x + y;                       // dependent DADD
            if (OP == 1) x = x * y;                       // dependent DMUL
            if (OP == 2) x = (x < y) ? x + 1.0 : y;       // DSETP + FSEL x2 + DADD
            if (OP == 3) { x = x + y; z = z + y; }        // 2 independent DADD chains
            if (OP == 4) { x = x + y; z = z + y; w = w + y; a = a + y; }  // 4 chains
            if (OP == 5) x = __fma_rn(x, y, b);           // DFMA
            if (OP == 6) x = (x < y) ? y : x;             // pure min/max select
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + z + w + a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int warps, int ops_per_iter) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 1024 * 4);
    cudaMallocManaged(&cyc, sizeof(long long));
    const int n = 2000;
    chain<OP><<<1, 32 * warps>>>(out, cyc, 1.0, 1.0000001, n);
    cudaDeviceSynchronize();
    chain<OP><<<1, 32 * warps>>>(out, cyc, 1.0, 1.0000001, n);
    cudaDeviceSynchronize();
    printf("%-28s warps/SM=%2d  cycles per op-group = %.2f\n", name, warps, (double)*cyc / (n * 16.0));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16}) {
        if (w == 1) { run<0>("DADD dep", 1, 1); run<1>("DMUL dep", 1, 1); run<5>("DFMA dep", 1, 1);
                      run<2>("DSETP+FSEL+DADD dep", 1, 1); run<6>("min/max select dep", 1, 1);
                      run<3>("2x DADD indep", 1, 2); run<4>("4x DADD indep", 1, 4); }
        if (w == 4) { run<0>("DADD dep", 4, 1); run<4>("4x DADD indep", 4, 4); run<6>("min/max select dep", 4, 1);}
        if (w == 8) { run<0>("DADD dep", 8, 1); run<4>("4x DADD indep", 8, 4); }
        if (w == 16) { run<0>("DADD dep", 16, 1); run<4>("4x DADD indep", 16, 4); run<2>("DSETP+FSEL+DADD dep", 16, 1);}
    }
    return 0;
}
