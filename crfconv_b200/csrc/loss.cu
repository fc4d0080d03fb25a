// Class-weighted cross entropy over [M, C] logits — the criterion of the training step (reference: trainval.py:66-70,100-104,
// torch.nn.CrossEntropyLoss(weight=…, ignore_index=…) on the [B·N, n_classes] output of PointConvResNet).  The library kernels
// behind F.cross_entropy cost 415 us per step at 245,760 points (a single-CTA nll_loss reduction forward and backward); here the
// forward is one pass (log-sum-exp, weighted loss, the two sums of the 'mean' reduction in double precision) and the backward
// recomputes the softmax from the logits: 13 MB per pass.
#include "../../include/crfconv_b200.h"
#include "common.cuh"

namespace crf {
namespace loss {

constexpr int kMaxC = 64;

template <int CMAX>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ tgt, const float* __restrict__ w,
                                                     int64_t M, int C, int64_t ignore, double* sums) {
    __shared__ double s_red[2][8];
    double ls = 0.0, ws = 0.0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = __ldg(tgt + m);
        if (t == ignore || t < 0 || t >= C) continue;
        const float* r = x + m * C;
        float v[CMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { v[c] = __ldg(r + c); mx = fmaxf(mx, v[c]); }
        float se = 0.f, xt = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { se += expf(v[c] - mx); if (c == (int)t) xt = v[c]; }
        const float wt = w ? __ldg(w + t) : 1.0f;
        ls += (double)(wt * (logf(se) + mx - xt));
        ws += (double)wt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ls += __shfl_xor_sync(0xffffffffu, ls, o); ws += __shfl_xor_sync(0xffffffffu, ws, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = ls; s_red[1][warp] = ws; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double tot = 0.0;
        for (int i = 0; i < 8; ++i) tot += s_red[threadIdx.x][i];
        atomicAdd(sums + threadIdx.x, tot);
    }
}

// dlogits[m, c] = g · w[t] · (softmax_c − [c == t]),  g = gout (sum) or gout / Σw (mean);  ignored rows get zeros
template <int CMAX>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ tgt, const float* __restrict__ w,
                                                     int64_t M, int C, int64_t ignore, const double* __restrict__ sums,
                                                     const float* __restrict__ gout, int mean, float* __restrict__ dx) {
    const float gs = mean ? (float)((double)__ldg(gout) / fmax(sums[1], 1e-300)) : __ldg(gout);
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = __ldg(tgt + m);
        float* d = dx + m * C;
        if (t == ignore || t < 0 || t >= C) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) d[c] = 0.f;
            continue;
        }
        const float* r = x + m * C;
        float v[CMAX];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { v[c] = __ldg(r + c); mx = fmaxf(mx, v[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { v[c] = expf(v[c] - mx); se += v[c]; }
        const float f = gs * (w ? __ldg(w + t) : 1.0f);
        const float inv = f / se;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) d[c] = fmaf(v[c], inv, c == (int)t ? -f : 0.f);
    }
}

}  // namespace loss
}  // namespace crf

using namespace crf;

extern "C" {

// sums[0] += Σ_m w[t_m]·(logsumexp(x_m) − x_m[t_m]),  sums[1] += Σ_m w[t_m]   over rows with t_m != ignore_index (doubles, zeroed by the caller).
// weight may be NULL (all ones).  C ≤ 64.
int crfconv_cross_entropy_fwd(const float* logits, const int64_t* target, const float* weight, int64_t M, int C, int64_t ignore_index,
                              double* sums, void* stream) {
    if (!logits || !target || !sums || M < 0 || C < 1) return CRF_ERR_INVALID_ARG;
    if (C > loss::kMaxC) return CRF_ERR_UNSUPPORTED;
    if (M == 0) return CRF_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(M, 256), (int64_t)kNumSMs * 8);
    if (C <= 16) loss::ce_fwd_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, weight, M, C, ignore_index, sums);
    else loss::ce_fwd_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, weight, M, C, ignore_index, sums);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

// dlogits = d loss / d logits for loss = sums[0] (mean == 0) or sums[0] / sums[1] (mean != 0), scaled by the device scalar *gout.
int crfconv_cross_entropy_bwd(const float* logits, const int64_t* target, const float* weight, int64_t M, int C, int64_t ignore_index,
                              const double* sums, const float* gout, int mean, float* dlogits, void* stream) {
    if (!logits || !target || !sums || !gout || !dlogits || M < 0 || C < 1) return CRF_ERR_INVALID_ARG;
    if (C > loss::kMaxC) return CRF_ERR_UNSUPPORTED;
    if (M == 0) return CRF_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(M, 256), (int64_t)kNumSMs * 8);
    if (C <= 16) loss::ce_bwd_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, weight, M, C, ignore_index, sums, gout, mean, dlogits);
    else loss::ce_bwd_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, target, weight, M, C, ignore_index, sums, gout, mean, dlogits);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

}  // extern "C"
