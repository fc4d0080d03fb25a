// Graph builders and edge-list message passing for the PyG-API family of the reference (models/continuous_crf_conv.py,
// models/discrete_crf_conv.py, models/point_conv.py:140-195), whose graph construction lives in torch_cluster / torch_points_kernels
// (third party, not under /root/reference — semantics restated):
//   fps_kernel            farthest point sampling, one CTA per cloud (ragged clouds through CSR offsets)
//   edge_softmax_*        s_e = softmax over the edges of a target node of −‖y_row − y_col‖²  (torch_geometric.utils.softmax:
//                         exp(a − max_group) / (Σ_group + 1e-16)), forward and backward, edges grouped by target (CSR)
//   edge_gauss_*          w_e = Σ_k W_k·exp(−‖f_k[col] − f_k[row]‖²)                           (discrete_crf_conv.py:53-56)
//   spmm_*                out_i = Σ_{e: row_e = i} w_e·x[col_e]  (scatter_add of weighted messages), forward and backward
// One warp per target node; lanes stride over the channels (any channel count; 128-bit loads when it is a multiple of 4).
#include "../../include/crfconv_b200.h"
#include "common.cuh"

namespace crf {
namespace graph {

__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, float px, float py, float pz) {   // nanoflann order, no FMA (knn.cu)
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ------------------------------------------------------------------------------------------------ farthest point sampling
// Cloud b = points [ptr[b], ptr[b+1]); selects nsample[b] points starting from local index start[b]; out[out_ptr[b] + s] = GLOBAL index
// of the s-th pick.  dist: workspace of ptr[B] floats.  Ties → smallest index (the block reduction compares (distance, −index)).
constexpr int kFpsThreads = 1024;
__global__ void __launch_bounds__(kFpsThreads) fps_kernel(const float* __restrict__ pos, const int64_t* __restrict__ ptr,
                                                          const int64_t* __restrict__ nsample, const int64_t* __restrict__ start,
                                                          float* __restrict__ dist_all, int64_t* __restrict__ out,
                                                          const int64_t* __restrict__ out_ptr) {
    __shared__ float s_d[kFpsThreads / 32];
    __shared__ int s_i[kFpsThreads / 32];
    __shared__ int s_cur;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = ptr[b];
    const int n = (int)(ptr[b + 1] - p0);
    const float* P = pos + 3 * p0;
    float* dist = dist_all + p0;
    for (int i = tid; i < n; i += kFpsThreads) dist[i] = INFINITY;
    if (tid == 0) s_cur = (int)start[b];
    __syncthreads();
    const int ns = (int)nsample[b];
    for (int s = 0; s < ns; ++s) {
        const int cur = s_cur;
        if (tid == 0) out[out_ptr[b] + s] = p0 + cur;
        const float cx = __ldg(P + 3 * (size_t)cur), cy = __ldg(P + 3 * (size_t)cur + 1), cz = __ldg(P + 3 * (size_t)cur + 2);
        float best = -1.0f;
        int arg = 0x7fffffff;
        for (int i = tid; i < n; i += kFpsThreads) {
            const float d = fminf(dist[i], sqdist3(cx, cy, cz, __ldg(P + 3 * (size_t)i), __ldg(P + 3 * (size_t)i + 1), __ldg(P + 3 * (size_t)i + 2)));
            dist[i] = d;
            if (d > best) { best = d; arg = i; }                  // i ascends per thread: the first maximum keeps the smallest index
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
            if (od > best || (od == best && oi < arg)) { best = od; arg = oi; }
        }
        if (lane == 0) { s_d[warp] = best; s_i[warp] = arg; }
        __syncthreads();
        if (warp == 0) {
            best = s_d[lane]; arg = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, arg, o);
                if (od > best || (od == best && oi < arg)) { best = od; arg = oi; }
            }
            if (lane == 0) s_cur = arg;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ edge softmax of −‖Δy‖²
// Edges are grouped by target: node i owns edges [eptr[i], eptr[i+1]); col[e] = source node.  y [N, C].
__global__ void __launch_bounds__(256) edge_softmax_fwd_kernel(const float* __restrict__ y, const int64_t* __restrict__ eptr,
                                                               const int64_t* __restrict__ col, float* __restrict__ s, int64_t N, int C) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const int64_t e0 = eptr[i], e1 = eptr[i + 1];
    const float* yi = y + i * C;
    float mx = -INFINITY;
    for (int64_t e = e0; e < e1; ++e) {                          // pass 1: a_e = −‖y_i − y_j‖², kept in s
        const float* yj = y + col[e] * C;
        float d = 0.f;
        for (int c = lane; c < C; c += 32) { const float t = __ldg(yi + c) - __ldg(yj + c); d = fmaf(t, t, d); }
        d = warp_sum(d);
        if (lane == 0) s[e] = -d;
        mx = fmaxf(mx, -d);
    }
    __syncwarp();
    float sum = 0.f;
    for (int64_t e = e0 + lane; e < e1; e += 32) sum += __expf(s[e] - mx);
    sum = warp_sum(sum) + 1e-16f;                                // torch_geometric.utils.softmax
    for (int64_t e = e0 + lane; e < e1; e += 32) s[e] = __expf(s[e] - mx) / sum;
}

// Given ds (gradient wrt s): da_e = s_e·(ds_e − Σ_k s_k ds_k)·(Σ/(Σ+1e-16) ≈ 1);  a_e = −‖y_i − y_j‖²  ⇒
// dy_i += −2·da_e·(y_i − y_j),  dy_j += +2·da_e·(y_i − y_j)      (dy zero-initialised by the caller; atomics on the source rows)
__global__ void __launch_bounds__(256) edge_softmax_bwd_kernel(const float* __restrict__ y, const int64_t* __restrict__ eptr,
                                                               const int64_t* __restrict__ col, const float* __restrict__ s,
                                                               const float* __restrict__ ds, float* dy, int64_t N, int C) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const int64_t e0 = eptr[i], e1 = eptr[i + 1];
    float dot = 0.f;
    for (int64_t e = e0 + lane; e < e1; e += 32) dot = fmaf(s[e], ds[e], dot);
    dot = warp_sum(dot);
    const float* yi = y + i * C;
    for (int64_t e = e0; e < e1; ++e) {
        const float da2 = 2.0f * s[e] * (ds[e] - dot);
        const int64_t j = col[e];
        const float* yj = y + j * C;
        for (int c = lane; c < C; c += 32) {
            const float g = da2 * (__ldg(yi + c) - __ldg(yj + c));
            atomicAdd(dy + i * C + c, -g);
            atomicAdd(dy + j * C + c, g);
        }
    }
}

// ------------------------------------------------------------------------------------------------ multi-kernel Gaussian edge weights
// f [N, Kk·H] (kernel-major blocks of H), Wk [Kk]:  w_e = Σ_k Wk[k]·exp(−‖f_k[col_e] − f_k[row_e]‖²);  g [E, Kk] keeps exp(·) for the backward
__global__ void __launch_bounds__(256) edge_gauss_fwd_kernel(const float* __restrict__ f, const int64_t* __restrict__ eptr,
                                                             const int64_t* __restrict__ col, const float* __restrict__ Wk,
                                                             float* __restrict__ w, float* __restrict__ g, int64_t N, int Kk, int H) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const float* fi = f + i * Kk * H;
    for (int64_t e = eptr[i]; e < eptr[i + 1]; ++e) {
        const float* fj = f + col[e] * Kk * H;
        float acc = 0.f;
        for (int k = 0; k < Kk; ++k) {
            float d = 0.f;
            for (int c = lane; c < H; c += 32) { const float t = __ldg(fj + k * H + c) - __ldg(fi + k * H + c); d = fmaf(t, t, d); }
            d = warp_sum(d);
            const float ex = __expf(-d);
            if (lane == 0) g[e * Kk + k] = ex;
            acc = fmaf(__ldg(Wk + k), ex, acc);
        }
        if (lane == 0) w[e] = acc;
    }
}

// dw [E] → dWk[k] += Σ_e dw_e·g_ek;  df[col] += −2·dw_e·Wk[k]·g_ek·(f_k[col] − f_k[row]),  df[row] −= the same
__global__ void __launch_bounds__(256) edge_gauss_bwd_kernel(const float* __restrict__ f, const int64_t* __restrict__ eptr,
                                                             const int64_t* __restrict__ col, const float* __restrict__ Wk,
                                                             const float* __restrict__ g, const float* __restrict__ dw, float* df, float* dWk,
                                                             int64_t N, int Kk, int H) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const float* fi = f + i * Kk * H;
    for (int k = 0; k < Kk; ++k) {
        float dwk = 0.f;
        const float wk = __ldg(Wk + k);
        for (int64_t e = eptr[i]; e < eptr[i + 1]; ++e) {
            const int64_t j = col[e];
            const float gek = g[e * Kk + k], de = dw[e];
            if (lane == 0) dwk = fmaf(de, gek, dwk);
            const float coef = -2.0f * de * wk * gek;
            const float* fj = f + j * Kk * H;
            for (int c = lane; c < H; c += 32) {
                const float t = coef * (__ldg(fj + k * H + c) - __ldg(fi + k * H + c));
                atomicAdd(df + j * Kk * H + k * H + c, t);
                atomicAdd(df + i * Kk * H + k * H + c, -t);
            }
        }
        if (lane == 0 && dwk != 0.f) atomicAdd(dWk + k, dwk);
    }
}

// ------------------------------------------------------------------------------------------------ weighted aggregation
// out[i, :] = Σ_{e in node i} w_e·x[col_e, :]
__global__ void __launch_bounds__(256) spmm_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ eptr,
                                                       const int64_t* __restrict__ col, const float* __restrict__ w, float* __restrict__ out,
                                                       int64_t N, int C) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    const int64_t e0 = eptr[i], e1 = eptr[i + 1];
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.f;
        if (c < C)
            for (int64_t e = e0; e < e1; ++e) acc = fmaf(w[e], __ldg(x + col[e] * C + c), acc);
        if (c < C) out[i * C + c] = acc;
    }
}

// g [N, C] = gradient wrt out:  dw_e = ⟨g_i, x[col_e]⟩;  dx[col_e, :] += w_e·g_i   (dx zero-initialised; either output may be null)
__global__ void __launch_bounds__(256) spmm_bwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ eptr,
                                                       const int64_t* __restrict__ col, const float* __restrict__ w, const float* __restrict__ g,
                                                       float* dw, float* dx, int64_t N, int C) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= N) return;
    const int lane = threadIdx.x & 31;
    for (int64_t e = eptr[i]; e < eptr[i + 1]; ++e) {
        const int64_t j = col[e];
        const float we = w[e];
        float dot = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float gi = __ldg(g + i * C + c);
            dot = fmaf(gi, __ldg(x + j * C + c), dot);
            if (dx) atomicAdd(dx + j * C + c, we * gi);
        }
        if (dw) {
            dot = warp_sum(dot);
            if (lane == 0) dw[e] = dot;
        }
    }
}

}  // namespace graph
}  // namespace crf

using namespace crf;

extern "C" {

// Farthest point sampling (torch_points_kernels.furthest_point_sampling / torch_cluster.fps as used by
// datasets/s3dis_dataset.py:434-437 and models/point_conv.py:175).  pos [ptr[B], 3]; ptr [B+1] CSR offsets of the clouds; nsample,
// start [B]; out [out_ptr[B]] receives GLOBAL point indices, cloud after cloud; dist_ws: ptr[B] floats.  All device pointers.
int crfconv_fps(const float* pos, const int64_t* ptr, int64_t B, const int64_t* nsample, const int64_t* start, int64_t* out,
                const int64_t* out_ptr, float* dist_ws, void* stream) {
    if (B < 0 || !pos || !ptr || !nsample || !start || !out || !out_ptr || !dist_ws) return CRF_ERR_INVALID_ARG;
    if (B == 0) return CRF_OK;
    graph::fps_kernel<<<(unsigned)B, graph::kFpsThreads, 0, (cudaStream_t)stream>>>(pos, ptr, nsample, start, dist_ws, out, out_ptr);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

static inline unsigned warp_grid(int64_t N) { return (unsigned)ceil_div(N, 8); }

// s[E] = softmax over each target node's edges of −‖y_i − y_col‖² (continuous_crf_conv.py:55-56,115-116); edges grouped by target.
int crfconv_edge_softmax_fwd(const float* y, const int64_t* eptr, const int64_t* col, float* s, int64_t N, int C, void* stream) {
    if (N < 0 || C <= 0 || !y || !eptr || !col || !s) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::edge_softmax_fwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(y, eptr, col, s, N, C);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}
int crfconv_edge_softmax_bwd(const float* y, const int64_t* eptr, const int64_t* col, const float* s, const float* ds, float* dy, int64_t N,
                             int C, void* stream) {
    if (N < 0 || C <= 0 || !y || !eptr || !col || !s || !ds || !dy) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::edge_softmax_bwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(y, eptr, col, s, ds, dy, N, C);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}
// w[E] = Σ_k Wk[k]·exp(−‖f_k[col] − f_k[row]‖²), f [N, Kk·H] (discrete_crf_conv.py:49-56); g [E, Kk] saves the exponentials.
int crfconv_edge_gauss_fwd(const float* f, const int64_t* eptr, const int64_t* col, const float* Wk, float* w, float* g, int64_t N, int Kk,
                           int H, void* stream) {
    if (N < 0 || Kk <= 0 || H <= 0 || !f || !eptr || !col || !Wk || !w || !g) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::edge_gauss_fwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(f, eptr, col, Wk, w, g, N, Kk, H);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}
int crfconv_edge_gauss_bwd(const float* f, const int64_t* eptr, const int64_t* col, const float* Wk, const float* g, const float* dw,
                           float* df, float* dWk, int64_t N, int Kk, int H, void* stream) {
    if (N < 0 || Kk <= 0 || H <= 0 || !f || !eptr || !col || !Wk || !g || !dw || !df || !dWk) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::edge_gauss_bwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(f, eptr, col, Wk, g, dw, df, dWk, N, Kk, H);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}
// out[N,C] = scatter_add(w ⊙ x[col], row) (continuous_crf_conv.py:64-65, discrete_crf_conv.py:61) and its backward.
int crfconv_spmm_fwd(const float* x, const int64_t* eptr, const int64_t* col, const float* w, float* out, int64_t N, int C, void* stream) {
    if (N < 0 || C <= 0 || !x || !eptr || !col || !w || !out) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::spmm_fwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(x, eptr, col, w, out, N, C);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}
int crfconv_spmm_bwd(const float* x, const int64_t* eptr, const int64_t* col, const float* w, const float* g, float* dw, float* dx, int64_t N,
                     int C, void* stream) {
    if (N < 0 || C <= 0 || !x || !eptr || !col || !w || !g) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    graph::spmm_bwd_kernel<<<warp_grid(N), 256, 0, (cudaStream_t)stream>>>(x, eptr, col, w, g, dw, dx, N, C);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

}  // extern "C"
