// Microbenchmark of the two-level query sort: K2 appends 32-byte records to one frontier per tile
// (6859 tiles, global atomics), K3 streams that array and scatters each record to its (tile, x, y)
// bin inside the tile's own region (196 bins per tile, global atomics on 1.3 M cursors).
// One record per thread, grid in array order (the launch shape that reached streaming speed in
// windowed_scatter.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct __align__(32) Rec { double v[4]; };
struct Pt { double x, y, z; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t pick(uint32_t h, uint32_t bins) {
    return static_cast<uint32_t>((static_cast<uint64_t>(h) * bins) >> 32);
}

template <int PER>
__global__ void k2_tiles(const Pt* __restrict__ in, Rec* __restrict__ out, unsigned* cursor, long long n, uint32_t tiles,
                         long long cap) {
    const long long i0 = (blockIdx.x * (long long)blockDim.x) * PER + threadIdx.x;
    Pt p[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) { const long long i = i0 + u * blockDim.x; if (i < n) p[u] = in[i]; }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + u * blockDim.x;
        if (i >= n) continue;
        const uint32_t t = pick(hash32(static_cast<uint32_t>(i)), tiles);
        const unsigned pos = atomicAdd(cursor + t, 1u);
        Rec r; r.v[0] = p[u].x; r.v[1] = p[u].y; r.v[2] = p[u].z; r.v[3] = __longlong_as_double(i);
        out[t * cap + pos] = r;
    }
}

template <int PER>
__global__ void k3_fine(const Rec* __restrict__ in, Rec* __restrict__ out, unsigned* cursor, long long n, long long cap,
                        uint32_t fine, long long fcap) {
    const long long i0 = (blockIdx.x * (long long)blockDim.x) * PER + threadIdx.x;
    Rec r[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) { const long long i = i0 + u * blockDim.x; if (i < n) r[u] = in[i]; }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const long long i = i0 + u * blockDim.x;
        if (i >= n) continue;
        const long long t = i / cap;
        const uint32_t f = pick(hash32(static_cast<uint32_t>(__double_as_longlong(r[u].v[3]))), fine);
        const unsigned pos = atomicAdd(cursor + t * fine + f, 1u);
        if (pos < fcap) out[t * cap + f * fcap + pos] = r[u];
    }
}

template <typename F>
float time_ms(F&& f, cudaEvent_t e0, cudaEvent_t e1) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    const long long n = 1ll << 26;
    const uint32_t tiles = 6859, fine = 196;
    const long long cap = (n + (n >> 3)) / tiles, fcap = cap / fine * 2;
    Pt* pts; Rec *a, *b; unsigned *cur, *fcur;
    cudaMalloc(&pts, n * sizeof(Pt)); cudaMalloc(&a, tiles * cap * sizeof(Rec)); cudaMalloc(&b, tiles * cap * sizeof(Rec) * 2);
    cudaMalloc(&cur, sizeof(unsigned) * tiles); cudaMalloc(&fcur, sizeof(unsigned) * tiles * fine);
    cudaMemset(pts, 0, n * sizeof(Pt)); cudaMemset(a, 0, tiles * cap * sizeof(Rec));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long n3 = tiles * cap;   // K3 streams the whole (slack included) tile-major array
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(cur, 0, sizeof(unsigned) * tiles);
        float a1 = time_ms([&] { k2_tiles<1><<<(unsigned)((n + 255) / 256), 256>>>(pts, a, cur, n, tiles, cap); }, e0, e1);
        cudaMemset(cur, 0, sizeof(unsigned) * tiles);
        float a4 = time_ms([&] { k2_tiles<4><<<(unsigned)((n + 1023) / 1024), 256>>>(pts, a, cur, n, tiles, cap); }, e0, e1);
        cudaMemset(fcur, 0, sizeof(unsigned) * tiles * fine);
        float b1 = time_ms([&] { k3_fine<1><<<(unsigned)((n3 + 255) / 256), 256>>>(a, b, fcur, n3, cap, fine, fcap); }, e0, e1);
        cudaMemset(fcur, 0, sizeof(unsigned) * tiles * fine);
        float b4 = time_ms([&] { k3_fine<4><<<(unsigned)((n3 + 1023) / 1024), 256>>>(a, b, fcur, n3, cap, fine, fcap); }, e0, e1);
        printf("K2 (tile frontiers) PER=1 %.3f ms  PER=4 %.3f ms | K3 (fine, in-tile window, %lld recs) PER=1 %.3f ms  PER=4 %.3f ms\n",
               a1, a4, n3, b1, b4);
    }
    printf("err: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
