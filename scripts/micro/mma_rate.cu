// Microbenchmark: mma.sync throughput on sm_100a (tf32 m16n8k8, bf16 m16n8k16), per SM and chip-wide.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u}, b[2] = {threadIdx.x + 5u, 13u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int macs_per_mma) {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    for (int warps : {4, 8, 16, 32}) {
        int iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148 * 2, warps * 16>>>(out, 100);
        cudaEventRecord(e0);
        k<MODE><<<148 * 2, warps * 16>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = 148.0 * 2 * (warps / 2) * iters * 8;
        printf("%s warps/SM=%d: %.1f TFLOP/s dense (%.0f mma/us/SM)\n", name, warps, mmas * macs_per_mma * 2 / ms / 1e9, mmas / 148 / ms / 1e3);
    }
}
int main() { run<0>("tf32 m16n8k8 ", 16 * 8 * 8); run<1>("bf16 m16n8k16", 16 * 8 * 16); return 0; }
