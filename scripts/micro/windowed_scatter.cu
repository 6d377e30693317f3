// Microbenchmark: throughput of scattering 32-byte records when every block of 2^k consecutive
// source records lands in its own 2^k-record destination window (k = 26: fully random).
// Decides whether a two-level query sort (coarse bins, then an L2-resident fine scatter) can beat
// the random-sector bound.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o wscatter windowed_scatter.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) Rec { double v[4]; };

__device__ __forceinline__ uint32_t mix(uint32_t x, int k) {
    const uint32_t mask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    const int h = (k + 1) / 2;
    x = (x * 0x9E3779B1u) & mask; x ^= x >> h;
    x = (x * 0x85EBCA6Bu | 0) & mask; x ^= x >> h;
    x = (x * 0xC2B2AE35u) & mask; x ^= x >> h;
    return x & mask;
}

__global__ void scatter(const Rec* __restrict__ in, Rec* __restrict__ out, long long n, int k) {
    const uint32_t mask = (1u << k) - 1u;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t lo = static_cast<uint32_t>(i) & mask;
        const long long dst = (i & ~static_cast<long long>(mask)) | mix(lo, k);
        const Rec r = in[i];
        out[dst] = r;
    }
}
__global__ void gather(const Rec* __restrict__ in, Rec* __restrict__ out, long long n, int k) {
    const uint32_t mask = (1u << k) - 1u;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t lo = static_cast<uint32_t>(i) & mask;
        const long long src = (i & ~static_cast<long long>(mask)) | mix(lo, k);
        out[i] = in[src];
    }
}

int main() {
    const long long n = 1ll << 26;
    Rec *a, *b;
    cudaMalloc(&a, n * sizeof(Rec)); cudaMalloc(&b, n * sizeof(Rec));
    cudaMemset(a, 1, n * sizeof(Rec)); cudaMemset(b, 0, n * sizeof(Rec));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid_mode = 0; grid_mode < 2; ++grid_mode)
    for (int k = 12; k <= 26; k += 2) {
        // grid_mode 0: one pass over the array in order with a big grid (windows processed roughly one after another);
        // grid_mode 1: small persistent grid (148*8 CTAs) striding the array
        const int threads = 256;
        const int grid = grid_mode ? 148 * 8 : static_cast<int>(n / threads);
        float best_s = 1e9f, best_g = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0); scatter<<<grid, threads>>>(a, b, n, k); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best_s = ms < best_s ? ms : best_s;
            cudaEventRecord(e0); gather<<<grid, threads>>>(a, b, n, k); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); if (rep) best_g = ms < best_g ? ms : best_g;
        }
        printf("grid_mode %d window 2^%d records (%8.2f MB): scatter %.3f ms (%.1f Grec/s)  gather %.3f ms (%.1f Grec/s)\n",
               grid_mode, k, (double)(1ll << k) * 32 / 1e6, best_s, n / best_s / 1e6, best_g, n / best_g / 1e6);
    }
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
