// Dependent-chain latencies of the fp64 pipe on sm_100a (one warp, clock64 around N dependent ops).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_latency dp_latency.cu && ./dp_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(double* out, long long* cyc, double a, double b, int n) {
    double x = a + threadIdx.x * 1e-9;
    double y = b;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) { x = __fma_rn(x, y, b); }                       // DFMA
        if (MODE == 1) { x = __dsub_rn(b, __dmul_rn(x, y)); }           // DMUL -> DADD
        if (MODE == 2) { x = __ddiv_rn(b, x + 1.0); }                   // full division (+ add)
        if (MODE == 3) { x = __dadd_rn(x, y); }                         // DADD
        if (MODE == 4) { x = __dmul_rn(x, y); }                         // DMUL
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
    const int n = 4096;
    const char* names[] = {"DFMA", "DMUL->DADD", "__ddiv_rn(+DADD)", "DADD", "DMUL"};
    for (int warps = 1; warps <= 16; warps *= 2) {
        for (int mode = 0; mode < 5; ++mode) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: chain<0><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n); break;
                    case 1: chain<1><<<1, 32 * warps>>>(out, cyc, 1.0, 0.5, n); break;
                    case 2: chain<2><<<1, 32 * warps>>>(out, cyc, 1.0, 0.5, n); break;
                    case 3: chain<3><<<1, 32 * warps>>>(out, cyc, 1.0, 1e-9, n); break;
                    case 4: chain<4><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999999, n); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d  %-18s %.1f cycles per iteration\n", warps, names[mode], double(h) / n);
        }
    }
    return 0;
}
