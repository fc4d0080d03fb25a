// Neighbour gather / aggregation kernels of the dense-API point convolution on sm_100a.
//
// Replaces, without materialising the reference's [B,N',K,F] neighbour tensors or its [B,N'·K,F] int64 repeated index:
//   PointConv.gather_neighbors + (w * x).sum(2)          models/point_conv_big.py:25-35,56-57
//   PointConv._compute_weights' relative positions        models/point_conv_big.py:37-41   (centre − neighbour)
//   ResNetBBlock.max_pooling                               models/point_conv_big.py:74-77
//   Upsampling.upsampling backward (row scatter)           models/point_conv_big.py:97-101
// Feature rows are read with 128-bit loads (C/4 lanes per row); the edge-weight tensor H2 [B·N'·K, C] is consumed as the
// pre-BatchNorm output of weight_nn's second Linear with the BN affine applied on the fly.  The backward of the gather is
// a scatter along the transposed graph (red.global.add.v4.f32 into L2-resident rows).
#include <algorithm>

#include "../../include/crfconv_b200.h"
#include "common.cuh"

namespace crf {
namespace pc {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// rel[e, :] = centre[b, i, :] − support[b, idx[b, i, k], :],  e = (b·Nq + i)·K + k
__global__ void __launch_bounds__(256) relpos_kernel(const float* __restrict__ support, const float* __restrict__ centres,
                                                     const int64_t* __restrict__ idx, float* __restrict__ rel, int64_t E, int64_t Ns,
                                                     int64_t Nq, int K) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t row = e / K;               // b·Nq + i
    const int64_t b = row / Nq;
    const float* c = centres + row * 3;
    const float* s = support + (b * Ns + __ldg(idx + e)) * 3;
    rel[e * 3 + 0] = __ldg(c) - __ldg(s);
    rel[e * 3 + 1] = __ldg(c + 1) - __ldg(s + 1);
    rel[e * 3 + 2] = __ldg(c + 2) - __ldg(s + 2);
}

// out[row, c] = Σ_k (H2[e, c]·scale[c] + shift[c]) · x[b·Ns + idx[e], c]
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ H2,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const int64_t* __restrict__ idx, float* __restrict__ out, int64_t rows,
                                                            int64_t Ns, int64_t Nq, int K, int C4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C4) return;
    const int64_t row = t / C4;
    const int c = (int)(t % C4) * 4, C = C4 * 4;
    const int64_t base = (row / Nq) * Ns;
    const float4 sc = ld4(scale + c), sh = ld4(shift + c);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
        const int64_t e = row * K + k;
        const float4 h = ld4(H2 + e * C + c);
        const float4 xv = ld4(x + (base + __ldg(idx + e)) * C + c);
        acc.x = fmaf(fmaf(h.x, sc.x, sh.x), xv.x, acc.x);
        acc.y = fmaf(fmaf(h.y, sc.y, sh.y), xv.y, acc.y);
        acc.z = fmaf(fmaf(h.z, sc.z, sh.z), xv.z, acc.z);
        acc.w = fmaf(fmaf(h.w, sc.w, sh.w), xv.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + row * C + c) = acc;
}

// dWgt[e, c] = g[row, c]·x[src, c]   (gradient wrt the BatchNorm'd edge weight);   dx[src, c] += w[e, c]·g[row, c]
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ H2,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const int64_t* __restrict__ idx, const float* __restrict__ g,
                                                            float* __restrict__ dWgt, float* dx, int64_t rows, int64_t Ns, int64_t Nq,
                                                            int K, int C4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C4) return;
    const int64_t row = t / C4;
    const int c = (int)(t % C4) * 4, C = C4 * 4;
    const int64_t base = (row / Nq) * Ns;
    const float4 sc = ld4(scale + c), sh = ld4(shift + c);
    const float4 gv = ld4(g + row * C + c);
    for (int k = 0; k < K; ++k) {
        const int64_t e = row * K + k;
        const int64_t src = base + __ldg(idx + e);
        const float4 h = ld4(H2 + e * C + c);
        const float4 xv = ld4(x + src * C + c);
        *reinterpret_cast<float4*>(dWgt + e * C + c) = make_float4(gv.x * xv.x, gv.y * xv.y, gv.z * xv.z, gv.w * xv.w);
        if (dx)
            red_add_v4(dx + src * C + c, make_float4(fmaf(h.x, sc.x, sh.x) * gv.x, fmaf(h.y, sc.y, sh.y) * gv.y,
                                                      fmaf(h.z, sc.z, sh.z) * gv.z, fmaf(h.w, sc.w, sh.w) * gv.w));
    }
}

// out[row, c] = max_k x[b·Ns + idx[row, k], c];  arg[row, c] = source row of the (first) maximum
__global__ void __launch_bounds__(256) gather_max_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx,
                                                             float* __restrict__ out, int* __restrict__ arg, int64_t rows, int64_t Ns,
                                                             int64_t Nq, int K, int C4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C4) return;
    const int64_t row = t / C4;
    const int c = (int)(t % C4) * 4, C = C4 * 4;
    const int64_t base = (row / Nq) * Ns;
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int a[4] = {0, 0, 0, 0};
    for (int k = 0; k < K; ++k) {
        const int64_t src = base + __ldg(idx + row * K + k);
        const float4 v = ld4(x + src * C + c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (vv[u] > m[u] || k == 0) { m[u] = vv[u]; a[u] = (int)src; }   // first maximum wins, like torch.max
    }
    *reinterpret_cast<float4*>(out + row * C + c) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<int4*>(arg + row * C + c) = make_int4(a[0], a[1], a[2], a[3]);
}

// dx[arg[row, c], c] += g[row, c]
__global__ void __launch_bounds__(256) gather_max_bwd_kernel(const float* __restrict__ g, const int* __restrict__ arg, float* dx,
                                                             int64_t total, int C) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % C);
    atomicAdd(dx + (int64_t)arg[t] * C + c, __ldg(g + t));
}

// dS = g · (out > 0 ? 1 : slope)
__global__ void __launch_bounds__(256) lrelu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out, float slope,
                                                        float* __restrict__ dS, int64_t total4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 gv = ld4(g + 4 * i), o = ld4(out + 4 * i);
        reinterpret_cast<float4*>(dS)[i] = make_float4(o.x > 0.f ? gv.x : gv.x * slope, o.y > 0.f ? gv.y : gv.y * slope,
                                                       o.z > 0.f ? gv.z : gv.z * slope, o.w > 0.f ? gv.w : gv.w * slope);
    }
}

// y += x
__global__ void __launch_bounds__(256) add_inplace_kernel(float* y, const float* __restrict__ x, int64_t total4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = reinterpret_cast<float4*>(y)[i];
        const float4 b = ld4(x + 4 * i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(y)[i] = a;
    }
}

// dst[b·Ns + idx[m], :] += src[m, :]     (row scatter-add, C % 4 == 0)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* dst,
                                                               int64_t rows, int64_t Nq, int64_t Ns, int C4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * C4) return;
    const int64_t m = t / C4;
    const int c = (int)(t % C4) * 4;
    red_add_v4(dst + ((m / Nq) * Ns + __ldg(idx + m)) * (C4 * 4) + c, ld4(src + m * (C4 * 4) + c));
}

inline unsigned blocks(int64_t n) { return (unsigned)ceil_div(n, 256); }

}  // namespace pc
}  // namespace crf

using namespace crf;

extern "C" {

int crfconv_relpos(const float* support, const float* centres, const int64_t* idx, float* rel, int64_t B, int64_t Ns, int64_t Nq, int K,
                   void* stream) {
    if (B < 0 || Ns <= 0 || Nq < 0 || K <= 0 || !support || !centres || !idx || !rel) return CRF_ERR_INVALID_ARG;
    const int64_t E = B * Nq * K;
    if (E == 0) return CRF_OK;
    pc::relpos_kernel<<<pc::blocks(E), 256, 0, (cudaStream_t)stream>>>(support, centres, idx, rel, E, Ns, Nq, K);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pointconv_aggregate_fwd(const float* x, const float* H2, const float* scale, const float* shift, const int64_t* idx, float* out,
                                    int64_t B, int64_t Ns, int64_t Nq, int K, int C, void* stream) {
    if (B < 0 || Ns <= 0 || Nq < 0 || K <= 0 || C <= 0 || (C & 3) || !x || !H2 || !scale || !shift || !idx || !out) return CRF_ERR_INVALID_ARG;
    const int64_t rows = B * Nq;
    if (rows == 0) return CRF_OK;
    pc::aggregate_fwd_kernel<<<pc::blocks(rows * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, H2, scale, shift, idx, out, rows, Ns, Nq, K, C / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pointconv_aggregate_bwd(const float* x, const float* H2, const float* scale, const float* shift, const int64_t* idx,
                                    const float* g, float* dWgt, float* dx, int64_t B, int64_t Ns, int64_t Nq, int K, int C, void* stream) {
    if (B < 0 || Ns <= 0 || Nq < 0 || K <= 0 || C <= 0 || (C & 3) || !x || !H2 || !scale || !shift || !idx || !g || !dWgt) return CRF_ERR_INVALID_ARG;
    const int64_t rows = B * Nq;
    if (rows == 0) return CRF_OK;
    pc::aggregate_bwd_kernel<<<pc::blocks(rows * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, H2, scale, shift, idx, g, dWgt, dx, rows, Ns, Nq, K,
                                                                                        C / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_gather_max_fwd(const float* x, const int64_t* idx, float* out, int32_t* arg, int64_t B, int64_t Ns, int64_t Nq, int K, int C,
                           void* stream) {
    if (B < 0 || Ns <= 0 || Nq < 0 || K <= 0 || C <= 0 || (C & 3) || !x || !idx || !out || !arg || B * Ns > 0x7fffffff) return CRF_ERR_INVALID_ARG;
    const int64_t rows = B * Nq;
    if (rows == 0) return CRF_OK;
    pc::gather_max_fwd_kernel<<<pc::blocks(rows * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, idx, out, arg, rows, Ns, Nq, K, C / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_gather_max_bwd(const float* g, const int32_t* arg, float* dx, int64_t rows, int C, void* stream) {
    if (rows < 0 || C <= 0 || !g || !arg || !dx) return CRF_ERR_INVALID_ARG;
    if (rows == 0) return CRF_OK;
    pc::gather_max_bwd_kernel<<<pc::blocks(rows * C), 256, 0, (cudaStream_t)stream>>>(g, arg, dx, rows * C, C);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_lrelu_bwd(const float* g, const float* out, float slope, float* dS, int64_t numel, void* stream) {
    if (numel < 0 || (numel & 3) || !g || !out || !dS) return CRF_ERR_INVALID_ARG;
    if (numel == 0) return CRF_OK;
    pc::lrelu_bwd_kernel<<<(unsigned)std::min<int64_t>(ceil_div(numel / 4, 256), (int64_t)kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(g, out, slope, dS, numel / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_add_inplace(float* y, const float* x, int64_t numel, void* stream) {
    if (numel < 0 || (numel & 3) || !y || !x) return CRF_ERR_INVALID_ARG;
    if (numel == 0) return CRF_OK;
    pc::add_inplace_kernel<<<(unsigned)std::min<int64_t>(ceil_div(numel / 4, 256), (int64_t)kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(y, x, numel / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_scatter_add_rows(const float* src, const int64_t* idx, float* dst, int64_t B, int64_t Nq, int64_t Ns, int C, void* stream) {
    if (B < 0 || Nq < 0 || Ns <= 0 || C <= 0 || (C & 3) || !src || !idx || !dst) return CRF_ERR_INVALID_ARG;
    const int64_t rows = B * Nq;
    if (rows == 0) return CRF_OK;
    pc::scatter_add_rows_kernel<<<pc::blocks(rows * (C / 4)), 256, 0, (cudaStream_t)stream>>>(src, idx, dst, rows, Nq, Ns, C / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

}  // extern "C"
