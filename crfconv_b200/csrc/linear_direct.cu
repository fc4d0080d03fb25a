// Weight gradients of NARROW layers over MANY rows:  dW[Cout, Ktot] += dHᵀ·[prologue(X1) | X2],  Cout·Ktot small, M large.
//
// The network's level-0 layers (245,760 points, models/point_conv_big.py:61-88,130-140) are Linear(6→8), (6→32), (8→32), (32→8),
// (32→16), (32→128), (128→13): their weight gradients read 40-160 floats per row and produce a few hundred numbers.  The tiled
// kernels (linear.cu wgrad_kernel: 64×64 output tile, unpipelined shared-memory staging; linear2.cu wgrad2_kernel: needs ≥ 8
// warps' worth of output) ran them at 0.3-1.3 TB/s — 125-250 us each, 1.7 ms of an 11 ms step.  Here nothing is staged: both MMA
// operands of the contraction over rows come straight from global memory in fragment order (4-byte "transposed" loads: the 8 lanes
// of a fragment column read one 32-byte sector, every element is fetched exactly once), the BatchNorm-backward transform
//     dH = sc·(dV − k1 − (H − mu)·istd·k2),   dV = dY·lrelu'(sc·H + sh)
// and the input prologue lrelu(x·s1 + t1) are applied in registers, and each warp owns whole 16-row tiles, so there is no
// CTA-wide barrier in the stream.  3xTF32 (fp32-grade).  Same scheme as cl::in16_wgrad_kernel (crf_fused.cu), for any
// Cout ≤ 128 (MB blocks of 16) and Ktot ≤ 128 (NB blocks of 8) with MB·NB ≤ 32 accumulator blocks.
#include <algorithm>

#include "../../include/crfconv_b200.h"
#include "common.cuh"
#include "fused_common.cuh"
#include "linear_args.cuh"

namespace crf {
namespace lin {
namespace dw {

constexpr int kThreads = 256, kWarps = 8;

template <int MB, int NB>
__global__ void __launch_bounds__(kThreads, MB * NB > 4 ? 1 : 2) wgrad_direct_kernel(const WgradArgs a) {
    constexpr int CO = 16 * MB, KP = 8 * NB;
    constexpr bool TWO = (8 * MB + 2 * NB) <= 40;              // both 8-row k-steps of a tile loaded before the first MMA
    constexpr int KSB = TWO ? 2 : 1;
    __shared__ float s_w[CO * KP];
    __shared__ float4 s_par[CO];                                // (sc, sh, −sc·istd·k2, −sc·k1 + sc·istd·k2·mu)
    __shared__ float2 s_pro[KP];
    __shared__ float s_b[CO];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int C = a.Cout, Ktot = a.C1 + a.C2;
    const bool plain = a.bn.scale == nullptr;
    for (int i = tid; i < CO * KP; i += kThreads) s_w[i] = 0.f;
    for (int k = tid; k < CO; k += kThreads) {
        float4 pr = make_float4(1.f, 0.f, 0.f, 0.f);
        if (!plain && k < C) {
            const float sc = __ldg(a.bn.scale + k), sh = __ldg(a.bn.shift + k), mu = __ldg(a.bn.mean + k), is = __ldg(a.bn.invstd + k);
            const float k1 = __ldg(a.bn.k1 + k), k2 = __ldg(a.bn.k2 + k);
            pr = make_float4(sc, sh, -sc * is * k2, -sc * k1 + sc * is * k2 * mu);
        }
        s_par[k] = pr;
        s_b[k] = 0.f;
    }
    for (int k = tid; k < KP; k += kThreads) {
        float2 pr = make_float2(1.f, 0.f);
        if (a.scale1 && k < a.C1) pr = make_float2(__ldg(a.scale1 + k), __ldg(a.shift1 + k));
        s_pro[k] = pr;
    }
    __syncthreads();
    const float slope = a.bn.slope, slope1 = a.scale1 ? a.slope1 : 1.0f;
    const bool pro = a.scale1 != nullptr;
    float acc[MB][NB][4];
    float bs[MB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        bs[mb][0] = bs[mb][1] = 0.f;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mb][nb][e] = 0.f;
    }
    const int64_t ntiles = (a.M + 15) >> 4;
    for (int64_t tile = (int64_t)blockIdx.x * kWarps + warp; tile < ntiles; tile += (int64_t)gridDim.x * kWarps) {
#pragma unroll
        for (int k0 = 0; k0 < 2; k0 += KSB) {
            float tv[KSB][MB][4], th[KSB][MB][4], xa[KSB][NB], xb[KSB][NB];
            bool oka[KSB], okb[KSB];
#pragma unroll
            for (int kk = 0; kk < KSB; ++kk) {
                const int64_t ra = tile * 16 + 8 * (k0 + kk) + t, rb = ra + 4;
                oka[kk] = ra < a.M; okb[kk] = rb < a.M;
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                // fragment order: (co, ra), (co + 8, ra), (co, rb), (co + 8, rb)
                        const int co = 16 * mb + g + 8 * (q & 1);
                        const bool ok = ((q & 2) ? okb[kk] : oka[kk]) && co < C;
                        const int64_t off = ((q & 2) ? rb : ra) * C + co;
                        tv[kk][mb][q] = ok ? __ldg(a.dY + off) : 0.f;
                        th[kk][mb][q] = (ok && !plain) ? __ldg(a.H + off) : 0.f;
                    }
                int64_t sa = ra, sb = rb;
                if (a.idx1) {                                   // segment 1 rows gathered: src = (m / rows_dst)·rows_src + idx1[m]
                    if (oka[kk]) sa = (ra / a.rows_dst) * a.rows_src + __ldg(a.idx1 + ra);
                    if (okb[kk]) sb = (rb / a.rows_dst) * a.rows_src + __ldg(a.idx1 + rb);
                }
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const int col = nb * 8 + g;
                    float va = 0.f, vb = 0.f;
                    if (col < a.C1) {
                        if (oka[kk]) va = __ldg(a.X1 + sa * a.C1 + col);
                        if (okb[kk]) vb = __ldg(a.X1 + sb * a.C1 + col);
                    } else if (col < Ktot) {
                        if (oka[kk]) va = __ldg(a.X2 + ra * a.C2 + (col - a.C1));
                        if (okb[kk]) vb = __ldg(a.X2 + rb * a.C2 + (col - a.C1));
                    }
                    xa[kk][nb] = va; xb[kk][nb] = vb;
                }
            }
#pragma unroll
            for (int kk = 0; kk < KSB; ++kk) {
                cl::FragB b[NB];
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    float va = xa[kk][nb], vb = xb[kk][nb];
                    if (pro) {
                        const int col = nb * 8 + g;
                        const float2 pr = s_pro[col];
                        const float sl = col < a.C1 ? slope1 : 1.0f;
                        va = (oka[kk] && col < Ktot) ? cl::lrelu(fmaf(va, pr.x, pr.y), sl) : 0.f;
                        vb = (okb[kk] && col < Ktot) ? cl::lrelu(fmaf(vb, pr.x, pr.y), sl) : 0.f;
                    }
                    cl::make_b(b[nb], va, vb);
                }
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    float av[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float d = tv[kk][mb][q];
                        if (!plain) {
                            const int co = 16 * mb + g + 8 * (q & 1);
                            const bool ok = ((q & 2) ? okb[kk] : oka[kk]) && co < C;
                            const float4 p = s_par[co];
                            const float h = th[kk][mb][q];
                            const float dv = fmaf(h, p.x, p.y) > 0.f ? d : d * slope;
                            d = ok ? fmaf(p.x, dv, fmaf(p.z, h, p.w)) : 0.f;
                        }
                        av[q] = d;
                    }
                    bs[mb][0] += av[0] + av[2];
                    bs[mb][1] += av[1] + av[3];
                    cl::FragA f;
                    cl::make_a(f, av[0], av[1], av[2], av[3]);
#pragma unroll
                    for (int nb = 0; nb < NB; ++nb) cl::mma3(acc[mb][nb], f, b[nb]);
                }
            }
        }
    }
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            float* w0 = &s_w[(16 * mb + g) * KP + nb * 8 + 2 * t];
            atomicAdd(w0, acc[mb][nb][0]);
            atomicAdd(w0 + 1, acc[mb][nb][1]);
            atomicAdd(w0 + 8 * KP, acc[mb][nb][2]);
            atomicAdd(w0 + 8 * KP + 1, acc[mb][nb][3]);
        }
        if (a.dbias) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float v = bs[mb][e];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (t == 0) atomicAdd(&s_b[16 * mb + g + 8 * e], v);
            }
        }
    }
    __syncthreads();
    float* dst = a.dW + a.slot_stride * (int64_t)(blockIdx.x % kGradSlots);
    for (int i = tid; i < CO * KP; i += kThreads) {
        const int co = i / KP, k = i % KP;
        if (co < C && k < Ktot) atomicAdd(dst + (int64_t)co * Ktot + k, s_w[i]);
    }
    if (a.dbias)
        for (int i = tid; i < C; i += kThreads) atomicAdd(a.dbias + i, s_b[i]);
}

template <int MB, int NB>
inline int launch(const WgradArgs& a, cudaStream_t st) {
    constexpr int per_sm = MB * NB > 4 ? 1 : 2;
    const int64_t tiles = ceil_div(a.M, 16);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(tiles, kWarps), (int64_t)kNumSMs * per_sm));
    wgrad_direct_kernel<MB, NB><<<grid, kThreads, 0, st>>>(a);
    const cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? CRF_OK : (int)e;
}

}  // namespace dw

bool try_wgrad_direct(const WgradArgs& a, cudaStream_t st, int* rc) {
    const int Ktot = a.C1 + a.C2;
    if (a.bn.scale && a.bn.act_ref) return false;               // saved-activation branch selection: tiled kernels
    if (a.Cout < 1 || a.Cout > 128 || Ktot < 1 || Ktot > 128 || a.C1 < 1 || a.M < 1) return false;
    const int mb = a.Cout <= 16 ? 1 : (a.Cout <= 32 ? 2 : (a.Cout <= 64 ? 4 : 8));
    const int nb = Ktot <= 8 ? 1 : (Ktot <= 16 ? 2 : (Ktot <= 32 ? 4 : (Ktot <= 64 ? 8 : 16)));
    if (mb * nb > 32) return false;
#define CRF_DW(m, n)                         \
    if (mb == m && nb == n) {                \
        *rc = dw::launch<m, n>(a, st);       \
        return true;                         \
    }
    CRF_DW(1, 1) CRF_DW(1, 2) CRF_DW(1, 4) CRF_DW(1, 8) CRF_DW(1, 16)
    CRF_DW(2, 1) CRF_DW(2, 2) CRF_DW(2, 4) CRF_DW(2, 8) CRF_DW(2, 16)
    CRF_DW(4, 1) CRF_DW(4, 2) CRF_DW(4, 4) CRF_DW(4, 8)
    CRF_DW(8, 1) CRF_DW(8, 2) CRF_DW(8, 4)
#undef CRF_DW
    return false;
}

}  // namespace lin
}  // namespace crf
