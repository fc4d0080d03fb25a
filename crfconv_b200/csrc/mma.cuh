// Warp-level tensor-core helpers (mma.sync m16n8k8, tf32 inputs, f32 accumulate) used by the skinny Linear
// contractions of the hot path (M = B·N rows ≫ Cin, Cout ≤ a few hundred).  Precision modes:
//   kTf32x3 — error-compensated "3xTF32": a = a_hi + a_lo, b = b_hi + b_lo, D += a_lo·b_hi + a_hi·b_lo + a_hi·b_hi
//             (≈ fp32 product accuracy; this is the default so that layer outputs/gradients match the fp32
//             reference to 1e-3 through 6 stacked BatchNorm'd layers);
//   kTf32x1 — single tf32 pass (≈ 5e-4 relative per product), for the throughput experiments in DESIGN.md.
// Fragment layouts (PTX ISA, m16n8k8 .tf32): g = lane>>2, t = lane&3
//   A (16x8, row):  a0=(g, t) a1=(g+8, t) a2=(g, t+4) a3=(g+8, t+4)
//   B (8x8,  col):  b0=(k=t, n=g) b1=(k=t+4, n=g)
//   C/D (16x8):     c0=(g, 2t) c1=(g, 2t+1) c2=(g+8, 2t) c3=(g+8, 2t+1)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace crf {

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// D += A·B with A, B given as f32 values already arranged in fragment order.
template <bool X3>
__device__ __forceinline__ void mma_f32in(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], bh[2];
    if constexpr (X3) {
        uint32_t al[4], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
        mma_tf32(d, al, bh);
        mma_tf32(d, ah, bl);
        mma_tf32(d, ah, bh);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ah[i] = to_tf32(a[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) bh[i] = to_tf32(b[i]);
        mma_tf32(d, ah, bh);
    }
}

// Pre-split fragments (A split once per k-step and reused across all n-tiles).
struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };

template <bool X3>
__device__ __forceinline__ void make_frag_a(FragA& f, const float (&a)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (X3) split_tf32(a[i], f.hi[i], f.lo[i]);
        else { f.hi[i] = to_tf32(a[i]); f.lo[i] = 0; }
    }
}
template <bool X3>
__device__ __forceinline__ void make_frag_b(FragB& f, float b0, float b1) {
    if constexpr (X3) { split_tf32(b0, f.hi[0], f.lo[0]); split_tf32(b1, f.hi[1], f.lo[1]); }
    else { f.hi[0] = to_tf32(b0); f.hi[1] = to_tf32(b1); f.lo[0] = f.lo[1] = 0; }
}
template <bool X3>
__device__ __forceinline__ void mma_frag(float (&d)[4], const FragA& a, const FragB& b) {
    if constexpr (X3) {
        mma_tf32(d, a.lo, b.hi);
        mma_tf32(d, a.hi, b.lo);
    }
    mma_tf32(d, a.hi, b.hi);
}

}  // namespace crf
