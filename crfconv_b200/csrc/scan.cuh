// Device-wide exclusive scan (int32) used by the grid-subsampling chain: block-local scan of 4096-element tiles,
// recursive scan of the tile totals, uniform add.  Pure HBM streaming; not a hot kernel.
#pragma once
#include "common.cuh"

namespace crf {
namespace scan {

constexpr int kThreads = 1024;
constexpr int kItems = 4;
constexpr int kTile = kThreads * kItems;

__global__ void __launch_bounds__(kThreads) tile_scan_kernel(const int* in, int* out,   /* may alias */
                                                             int* __restrict__ tile_tot, int64_t n) {
    __shared__ int warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    int v[kItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    int run = (w ? warp_tot[w - 1] : 0) + x - sum;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kThreads - 1 && tile_tot) tile_tot[blockIdx.x] = run;
}

__global__ void __launch_bounds__(kThreads) tile_add_kernel(int* __restrict__ out, const int* __restrict__ tile_off,
                                                            int64_t n) {
    const int add = tile_off[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) out[base + k] += add;
}

inline size_t workspace_ints(int64_t n) {
    size_t tot = 0;
    while (n > kTile) {
        n = ceil_div(n, kTile);
        tot += align_up((size_t)n * sizeof(int), 256) / sizeof(int);
    }
    return tot + 64;
}

// out may alias in.  ws must hold workspace_ints(n) ints.
inline int exclusive(const int* in, int* out, int64_t n, int* ws, cudaStream_t st) {
    if (n <= 0) return CRF_OK;
    const int64_t tiles = ceil_div(n, kTile);
    if (tiles == 1) {
        tile_scan_kernel<<<1, kThreads, 0, st>>>(in, out, nullptr, n);
    } else {
        int* tot = ws;
        int* rest = ws + align_up((size_t)tiles * sizeof(int), 256) / sizeof(int);
        tile_scan_kernel<<<(unsigned)tiles, kThreads, 0, st>>>(in, out, tot, n);
        int rc = exclusive(tot, tot, tiles, rest, st);
        if (rc != CRF_OK) return rc;
        tile_add_kernel<<<(unsigned)tiles, kThreads, 0, st>>>(out, tot, n);
    }
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

}  // namespace scan
}  // namespace crf
