// Micro-benchmark: issue throughput of fp64 / ALU / FMA-pipe instruction mixes with 4 warps per
// SM sub-partition and 4 independent chains per warp (B200).  Prints warp-instructions per
// cycle per sub-partition for each mix.   nvcc -arch=sm_100a -O3 mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MIX>
__global__ void k(double* out, long long* cyc, double a, double b, int n, int sel) {
    double x[4], y = b;
    int z[4];
    for (int q = 0; q < 4; ++q) { x[q] = a + q + threadIdx.x; z[q] = threadIdx.x + q; }
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (MIX == 0) x[q] = x[q] + y;                                  // 1 DADD
                if (MIX == 1) { x[q] = x[q] + y; z[q] = z[q] * sel + u; }        // DADD + IMAD
                if (MIX == 2) { x[q] = x[q] + y; x[q] = (z[q] > sel) ? x[q] : y; }  // DADD + 2 FSEL
                if (MIX == 3) { x[q] = x[q] + y; x[q] = (x[q] < y) ? x[q] : y; }  // DADD+DSETP+2FSEL
                if (MIX == 4) { z[q] = z[q] * sel + u; z[q] = z[q] ^ (z[q] >> 3); }  // IMAD+SHF+LOP
                if (MIX == 5) { x[q] = x[q] + y; z[q] = z[q] * sel + u; z[q] = z[q] ^ (z[q] >> 3); x[q] = (z[q] > sel) ? x[q] : y; }
            }
        }
    }
    long long t1 = clock64();
    double s = 0; int zi = 0;
    for (int q = 0; q < 4; ++q) { s += x[q]; zi += z[q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + zi;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MIX>
void run(const char* name, int instr_per_group) {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 512);
    cudaMallocManaged(&cyc, sizeof(long long));
    const int n = 2000;
    for (int rep = 0; rep < 2; ++rep) { k<MIX><<<1, 512>>>(out, cyc, 1.0, 1.0000001, n, 12345); cudaDeviceSynchronize(); }
    const double groups = (double)n * 8 * 4;          // per warp
    const double instr = groups * instr_per_group * 4;  // 4 warps per sub-partition
    printf("%-34s cycles/group/warp %.2f  ->  IPC per sub-partition %.2f (nominal %d instr/group)\n",
           name, (double)*cyc / groups, instr / (double)*cyc, instr_per_group);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("DADD", 1);
    run<1>("DADD + IMAD", 2);
    run<2>("DADD + FSELx2", 3);
    run<3>("DADD + DSETP + FSELx2", 4);
    run<4>("IMAD + SHF + LOP3", 3);
    run<5>("DADD+IMAD+SHF+LOP+ISETP+FSELx2", 7);
    return 0;
}
