// Read-only / copy HBM bandwidth probe: what a streaming kernel can reach on this B200 as a function of bytes in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu && ./read_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) read_kernel(const float4* __restrict__ x, size_t n4, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { const float4 v = __ldg(x + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 1234.5678f) out[0] = acc;
}
template <int U>
__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ x, float4* __restrict__ y, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) y[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) y[i] = __ldg(x + i);
}
// row-strided pattern of the mma fragment loads: a warp instruction covers 8 rows x 64 B of a [rows, 64] f32 matrix
__global__ void __launch_bounds__(256) frag_read_kernel(const float* __restrict__ x, size_t rows, float* out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (size_t tile = warp; tile * 16 + 15 < rows; tile += nwarps) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[2 * j] = __ldg(reinterpret_cast<const float4*>(x + (tile * 16 + g) * 64 + 16 * j + 4 * t));
            v[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(x + (tile * 16 + g + 8) * 64 + 16 * j + 4 * t));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 1234.5678f) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 10) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const size_t bytes = (size_t)1 << 30;     // 1 GiB source, 1 GiB destination: far beyond the 126 MB L2
    float4 *x, *y; float* out;
    cudaMalloc(&x, bytes); cudaMalloc(&y, bytes); cudaMalloc(&out, 4);
    cudaMemset(x, 0, bytes); cudaMemset(y, 0, bytes);
    const size_t n4 = bytes / 16;
    for (int mult : {2, 4, 8}) {
        const int grid = 148 * mult;
        printf("grid = 148 x %d CTAs of 256 threads\n", mult);
        printf("  read  U=1 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { read_kernel<1><<<grid, 256>>>(x, n4, out); }));
        printf("  read  U=4 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { read_kernel<4><<<grid, 256>>>(x, n4, out); }));
        printf("  read  U=8 : %7.1f GB/s\n", bytes / 1e6 / time_ms([&] { read_kernel<8><<<grid, 256>>>(x, n4, out); }));
        printf("  copy  U=4 : %7.1f GB/s (read + write)\n", 2 * bytes / 1e6 / time_ms([&] { copy_kernel<4><<<grid, 256>>>(x, y, n4); }));
        printf("  copy  U=8 : %7.1f GB/s (read + write)\n", 2 * bytes / 1e6 / time_ms([&] { copy_kernel<8><<<grid, 256>>>(x, y, n4); }));
        printf("  frag-pattern read (8 rows x 64 B per instruction, 8 loads in flight): %7.1f GB/s\n",
               bytes / 1e6 / time_ms([&] { frag_read_kernel<<<grid, 256>>>(reinterpret_cast<const float*>(x), bytes / 256, out); }));
    }
    // the same on a 63 MB array (one [245760, 64] f32 activation): includes launch + ramp + tail
    const size_t small = (size_t)245760 * 256;
    printf("63 MB array, grid 148 x 8: read U=8 %7.1f GB/s, copy U=8 %7.1f GB/s\n",
           small / 1e6 / time_ms([&] { read_kernel<8><<<148 * 8, 256>>>(x, small / 16, out); }, 30),
           2 * small / 1e6 / time_ms([&] { copy_kernel<8><<<148 * 8, 256>>>(x, y, small / 16); }, 30));
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
