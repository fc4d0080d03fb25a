// PointConv (models/point_conv_big.py:8-58) without per-edge tensors, hidden width d = 8 / 16 (levels 0 and 1 of PointConvResNet: 245,760
// points and 3.9 M edges at d = 8 — the [E, 8] tensors H1, H2, dWgt, dA1 of the layer-by-layer path are 126 MB each and cost 0.87 ms per
// PointConv — and 983 k edges at d = 16).
//
//   w_e = BN2( W2·lrelu(BN1(W1·r_e)) ),  r_e = c_i − s_j,  out_i = Σ_k w_ik ⊙ x_j          (BatchNorm over all E edges)
//
// The edge MLP is 3→8→8: recomputing it costs ≈100 FMAs per edge, storing it 64 B.  Every pass below re-derives what it needs
// from r_e (12 B per edge, written once by the relpos pass) and keeps only per-channel / per-matrix SUMS:
//   forward   relpos_moments   r_e, Σr, Σrrᵀ            ⇒ BN1 statistics analytically (h1 = W1·r is linear in r)
//             fwd              P_i = Σ_k h2 ⊙ x_j, Q_i = Σ_k x_j, Σh2, Σh2², Σa1, Σa1a1ᵀ   ⇒ BN2 statistics; out = sc2 ⊙ P + sh2 ⊙ Q
//   backward  bwd1             dw_e = g_i ⊙ x_j: Σdw, Σdw·ĥ2, Σ dw a1ᵀ;  dx_j += w_e ⊙ g_i (vector reductions)
//             bwd2             dh2 = BN2'(dw), dv1 = (W2ᵀ dh2) ⊙ lrelu'(·): Σdv1, Σdv1·ĥ1, Σ dv1 rᵀ
//             param_grads      dW2 and dW1 from the sums alone (the BatchNorm backward is affine in per-edge quantities):
//                 dW2[c,b] = sc2[c]·( Σdw[c]a1[b] − k1'[c]·Σa1[b] − k2'[c]·is2[c]·( (W2·Σa1a1ᵀ)[c,b] − mu2[c]·Σa1[b] ) )
//                 dW1[c,a] = sc1[c]·( Σdv1[c]r[a] − k1[c]·Σr[a]  − k2[c]·is1[c]·( (W1·Σrrᵀ)[c,a]   − mu1[c]·Σr[a]  ) )
// One thread owns 8 channels of one centre point (d / 8 adjacent lanes per point; h1 / a1 are computed redundantly, h2 and everything
// downstream for the own channels; the input-gradient pass exchanges dh2 by shuffle) and walks the point's K edges; plain fp32 FMAs, sums
// go to kStatSlots partial slots and are folded in double precision.  Positions carry no gradient (as in the layer-by-layer path).
#include <algorithm>

#include "../../include/crfconv_b200.h"
#include "common.cuh"

namespace crf {
namespace pcf {

constexpr int DC = 8;                      // channels per thread
constexpr int kThreads = 128;
constexpr int kMom = 9;                    // Σr (3) | Σrrᵀ upper triangle xx xy xz yy yz zz (6)

__device__ __forceinline__ float lrelu(float v, float s) { return v > 0.f ? v : v * s; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Sum of a per-thread array over all threads of the CTA that own the same channel block, into slot (blockIdx.x % kStatSlots) of a
// [kStatSlots][total] float buffer.  map(i, blk) = position of local element i of channel block blk in the global layout.
template <int LPP, int NV, typename Map>
__device__ __forceinline__ void block_sums_to_slot(float (&v)[NV], float* slots, int total, float* s_acc, Map map) {
    const int lane = threadIdx.x & 31, blk = lane % LPP;
    for (int i = threadIdx.x; i < total; i += kThreads) s_acc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o >= LPP; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane < LPP) atomicAdd(&s_acc[map(i, blk)], x);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += kThreads) atomicAdd(slots + (size_t)(blockIdx.x % kStatSlots) * total + i, s_acc[i]);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------- relative positions + moments
__global__ void __launch_bounds__(256) relpos_moments_kernel(const float* __restrict__ support, const float* __restrict__ centres,
                                                             const int64_t* __restrict__ idx, float* __restrict__ rel, float* mom, int64_t E,
                                                             int64_t Ns, int64_t Nq, int K) {
    __shared__ float s_red[8][kMom];
    float m[kMom];
#pragma unroll
    for (int i = 0; i < kMom; ++i) m[i] = 0.f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / K, b = row / Nq;
        const float* c = centres + row * 3;
        const float* s = support + (b * Ns + __ldg(idx + e)) * 3;
        const float x = __ldg(c) - __ldg(s), y = __ldg(c + 1) - __ldg(s + 1), z = __ldg(c + 2) - __ldg(s + 2);
        rel[e * 3 + 0] = x; rel[e * 3 + 1] = y; rel[e * 3 + 2] = z;
        m[0] += x; m[1] += y; m[2] += z;
        m[3] = fmaf(x, x, m[3]); m[4] = fmaf(x, y, m[4]); m[5] = fmaf(x, z, m[5]);
        m[6] = fmaf(y, y, m[6]); m[7] = fmaf(y, z, m[7]); m[8] = fmaf(z, z, m[8]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kMom; ++i) {
        float x = m[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < kMom) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_red[w][threadIdx.x];
        atomicAdd(mom + (size_t)(blockIdx.x % kStatSlots) * kMom + threadIdx.x, tot);
    }
}

__device__ __forceinline__ double fold_slots(const float* slots, int nv, int i) {
    // all 64 loads are issued before the first add (a rolled loop put an L2 round trip into every iteration: 17-21 us per call)
    float v[kStatSlots];
#pragma unroll
    for (int p = 0; p < kStatSlots; ++p) v[p] = __ldcg(slots + (size_t)p * nv + i);
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < kStatSlots; ++p) s += (double)v[p];
    return s;
}
// symmetric 3×3 from the 6-entry upper triangle
__device__ __forceinline__ double sym3(const double* t, int a, int b) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return t[lo == 0 ? hi : (lo == 1 ? 2 + hi : 5)];
}

// Σh1[c] = W1[c]·Σr,  Σh1²[c] = W1[c]ᵀ·(Σrrᵀ)·W1[c]   → slot 0 of a [kStatSlots][2D] BatchNorm statistics buffer (other slots stay zero)
__global__ void stats1_kernel(const float* __restrict__ mom, const float* __restrict__ W1, float* stats1, int D) {
    __shared__ double m[kMom];
    if (threadIdx.x < kMom) m[threadIdx.x] = fold_slots(mom, kMom, threadIdx.x);
    __syncthreads();
    const int c = threadIdx.x;
    if (c >= D) return;
    double s = 0.0, q = 0.0;
    for (int a = 0; a < 3; ++a) {
        s += (double)W1[c * 3 + a] * m[a];
        for (int b = 0; b < 3; ++b) q += (double)W1[c * 3 + a] * (double)W1[c * 3 + b] * sym3(m + 3, a, b);
    }
    stats1[c] = (float)s;
    stats1[D + c] = (float)q;
}

// ---------------------------------------------------------------------------------------------------- the edge MLP, recomputed
template <int D>
struct EdgeW {                              // shared-memory image of the weights and the BatchNorm-1 affine
    float4 W1[D];                           // (w0, w1, w2, 0)
    float4 W2[D][D / 4];                    // row c: W2[c][0..D)
    float sc1[D], sh1[D];
};
template <int D>
__device__ __forceinline__ void stage_edge_w(EdgeW<D>& s, const float* W1, const float* W2, const float* sc1, const float* sh1) {
    for (int i = threadIdx.x; i < D; i += kThreads) {
        s.W1[i] = make_float4(__ldg(W1 + i * 3), __ldg(W1 + i * 3 + 1), __ldg(W1 + i * 3 + 2), 0.f);
        s.sc1[i] = __ldg(sc1 + i); s.sh1[i] = __ldg(sh1 + i);
    }
    for (int i = threadIdx.x; i < D * D / 4; i += kThreads) s.W2[i / (D / 4)][i % (D / 4)] = ldg4(W2 + 4 * i);
}
// h1, pre1, a1 for all D channels; h2 for the thread's own channels c0 .. c0 + 7
template <int D>
__device__ __forceinline__ void edge_mlp(const EdgeW<D>& s, int c0, float rx, float ry, float rz, float slope1, float (&h1)[D],
                                         float (&pre1)[D], float (&a1)[D], float (&h2)[DC]) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float4 w = s.W1[c];
        h1[c] = fmaf(w.x, rx, fmaf(w.y, ry, w.z * rz));
        pre1[c] = fmaf(h1[c], s.sc1[c], s.sh1[c]);
        a1[c] = lrelu(pre1[c], slope1);
    }
#pragma unroll
    for (int c = 0; c < DC; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
            const float4 w = s.W2[c0 + c][q];
            acc = fmaf(w.x, a1[4 * q], fmaf(w.y, a1[4 * q + 1], fmaf(w.z, a1[4 * q + 2], fmaf(w.w, a1[4 * q + 3], acc))));
        }
        h2[c] = acc;
    }
}
// the thread's 8 channels of row `row` of a [rows, D] tensor
template <int D>
__device__ __forceinline__ void load_x8(const float* x, int64_t row, int c0, float (&v)[DC]) {
    const float4 a = ldg4(x + row * D + c0), b = ldg4(x + row * D + c0 + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// The K edges of a point are contiguous (192 B of relative positions, 128 B of indices for K = 16): four edges at a time come in with
// three 128-bit + two 128-bit loads per thread instead of sixteen 4- / 8-byte ones (the per-lane stride makes every load instruction
// cost 32 L1 wavefronts whatever its width).  Requires K % 4 == 0 (16-byte alignment of every group).
struct Edge4 { float r[12]; int j[4]; };
__device__ __forceinline__ void load_edge4(const float* rel, const int64_t* idx, int64_t e, Edge4& q) {
    const float4 a = ldg4(rel + 3 * e), b = ldg4(rel + 3 * e + 4), c = ldg4(rel + 3 * e + 8);
    q.r[0] = a.x; q.r[1] = a.y; q.r[2] = a.z; q.r[3] = a.w; q.r[4] = b.x; q.r[5] = b.y; q.r[6] = b.z; q.r[7] = b.w;
    q.r[8] = c.x; q.r[9] = c.y; q.r[10] = c.z; q.r[11] = c.w;
    const longlong2 i0 = __ldg(reinterpret_cast<const longlong2*>(idx + e)), i1 = __ldg(reinterpret_cast<const longlong2*>(idx + e + 2));
    q.j[0] = (int)i0.x; q.j[1] = (int)i0.y; q.j[2] = (int)i1.x; q.j[3] = (int)i1.y;
}
// for (k ...) body(rx, ry, rz, j): vector path when K % 4 == 0
template <typename F>
__device__ __forceinline__ void for_edges(const float* rel, const int64_t* idx, int64_t p, int K, F&& body) {
    if ((K & 3) == 0) {
        for (int k = 0; k < K; k += 4) {
            Edge4 q;
            load_edge4(rel, idx, p * K + k, q);
#pragma unroll 1
            for (int u = 0; u < 4; ++u) body(q.r[3 * u], q.r[3 * u + 1], q.r[3 * u + 2], (int64_t)q.j[u]);
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const int64_t e = p * K + k;
            body(__ldg(rel + 3 * e), __ldg(rel + 3 * e + 1), __ldg(rel + 3 * e + 2), __ldg(idx + e));
        }
    }
}

// Every kernel: thread t of the grid ↔ (point p = t / LPP, channel block blk = t % LPP).  The LPP threads of a point are adjacent lanes
// and run the same trip counts (rows·LPP is padded per warp by the `p < rows` clamp below: a clamped thread contributes nothing).
struct FwdArgs {
    const float* x; const float* rel; const int64_t* idx;
    const float* W1; const float* W2; const float* sc1; const float* sh1; float slope1;
    float* P; float* Q;                     // [rows, D]
    float* stats2;                          // [kStatSlots][2D]        Σh2 | Σh2²
    float* asum;                            // [kStatSlots][D + D·D]   Σa1 | Σa1a1ᵀ (d = 8: upper triangle only, rest zero)
    int64_t rows, Ns, Nq; int K;
};

template <int D>
__global__ void __launch_bounds__(kThreads, D == 8 ? 3 : 2) fwd_kernel(const FwdArgs a) {
    constexpr int LPP = D / DC;
    constexpr int NA = LPP == 1 ? DC + DC * (DC + 1) / 2 : DC + DC * D;     // Σa1 (own) | rows of Σa1a1ᵀ (own; d = 8: upper triangle)
    __shared__ EdgeW<D> sw;
    __shared__ float s_acc[D + D * D];
    stage_edge_w(sw, a.W1, a.W2, a.sc1, a.sh1);
    __syncthreads();
    float st[2 * DC], as[NA];
#pragma unroll
    for (int i = 0; i < 2 * DC; ++i) st[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NA; ++i) as[i] = 0.f;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    for (int64_t tt = (int64_t)blockIdx.x * kThreads + threadIdx.x; tt < a.rows * LPP; tt += nthr) {
        const int64_t p = tt / LPP;
        const int c0 = (int)(tt % LPP) * DC;
        const int64_t base = (p / a.Nq) * a.Ns;
        float P[DC], Q[DC];
#pragma unroll
        for (int c = 0; c < DC; ++c) P[c] = Q[c] = 0.f;
        for_edges(a.rel, a.idx, p, a.K, [&](float rx, float ry, float rz, int64_t j) {
            float xj[DC], h1[D], pre1[D], a1[D], h2[DC];
            load_x8<D>(a.x, base + j, c0, xj);
            edge_mlp<D>(sw, c0, rx, ry, rz, a.slope1, h1, pre1, a1, h2);
#pragma unroll
            for (int c = 0; c < DC; ++c) {
                P[c] = fmaf(h2[c], xj[c], P[c]);
                Q[c] += xj[c];
                st[c] += h2[c];
                st[DC + c] = fmaf(h2[c], h2[c], st[DC + c]);
            }
            if constexpr (LPP == 1) {
#pragma unroll
                for (int c = 0; c < DC; ++c) as[c] += a1[c];
                int t = DC;
#pragma unroll
                for (int b = 0; b < DC; ++b)
#pragma unroll
                    for (int c = b; c < DC; ++c) { as[t] = fmaf(a1[b], a1[c], as[t]); ++t; }
            } else {
#pragma unroll
                for (int b = 0; b < DC; ++b) {
                    const float ab = c0 == 0 ? a1[b] : a1[DC + b];           // LPP == 2: own rows are a1[c0 + b]
                    as[b] += ab;
#pragma unroll
                    for (int c = 0; c < D; ++c) as[DC + b * D + c] = fmaf(ab, a1[c], as[DC + b * D + c]);
                }
            }
        });
        float4* Pp = reinterpret_cast<float4*>(a.P + p * D + c0);
        float4* Qp = reinterpret_cast<float4*>(a.Q + p * D + c0);
        Pp[0] = make_float4(P[0], P[1], P[2], P[3]); Pp[1] = make_float4(P[4], P[5], P[6], P[7]);
        Qp[0] = make_float4(Q[0], Q[1], Q[2], Q[3]); Qp[1] = make_float4(Q[4], Q[5], Q[6], Q[7]);
    }
    block_sums_to_slot<LPP>(st, a.stats2, 2 * D, s_acc, [](int i, int blk) { return (i < DC ? 0 : D - DC) + blk * DC + i; });
    block_sums_to_slot<LPP>(as, a.asum, D + D * D, s_acc, [](int i, int blk) {
        if (i < DC) return blk * DC + i;
        if (LPP == 1) {                                            // i-th upper-triangle element ↔ (b, c)
            int t = i - DC, b = 0;
            while (t >= DC - b) { t -= DC - b; ++b; }
            return D + b * D + (b + t);
        }
        return D + (blk * DC + (i - DC) / D) * D + (i - DC) % D;
    });
}

// out = sc2 ⊙ P + sh2 ⊙ Q
__global__ void __launch_bounds__(256) out_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ sc2,
                                                  const float* __restrict__ sh2, float* __restrict__ out, int64_t total4, int D4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D4) * 4;
        const float4 p = ldg4(P + 4 * i), q = ldg4(Q + 4 * i), sc = ldg4(sc2 + c), sh = ldg4(sh2 + c);
        reinterpret_cast<float4*>(out)[i] = make_float4(fmaf(sc.x, p.x, sh.x * q.x), fmaf(sc.y, p.y, sh.y * q.y), fmaf(sc.z, p.z, sh.z * q.z),
                                                        fmaf(sc.w, p.w, sh.w * q.w));
    }
}

// ---------------------------------------------------------------------------------------------------- backward, pass 1
struct Bwd1Args {
    const float* x; const float* rel; const int64_t* idx; const float* g;
    const float* W1; const float* W2; const float* sc1; const float* sh1; float slope1;
    const float* sc2; const float* sh2; const float* mu2; const float* is2;
    float* dx;                              // [B·Ns, D] zero-initialised, or NULL
    float* sums2;                           // [kStatSlots][2D]   Σdw | Σdw·ĥ2
    float* mdw;                             // [kStatSlots][D·D]  Σ dw[c]·a1[b]
    int64_t rows, Ns, Nq; int K;
};

template <int D>
__global__ void __launch_bounds__(kThreads, D == 8 ? 3 : 2) bwd1_kernel(const Bwd1Args a) {
    constexpr int LPP = D / DC;
    __shared__ EdgeW<D> sw;
    __shared__ float s_c[4][D];             // sc2, sh2, mu2, is2
    __shared__ float s_acc[D * D];
    stage_edge_w(sw, a.W1, a.W2, a.sc1, a.sh1);
    if (threadIdx.x < D) {
        s_c[0][threadIdx.x] = __ldg(a.sc2 + threadIdx.x); s_c[1][threadIdx.x] = __ldg(a.sh2 + threadIdx.x);
        s_c[2][threadIdx.x] = __ldg(a.mu2 + threadIdx.x); s_c[3][threadIdx.x] = __ldg(a.is2 + threadIdx.x);
    }
    __syncthreads();
    float sm[2 * DC], md[DC * D];
#pragma unroll
    for (int i = 0; i < 2 * DC; ++i) sm[i] = 0.f;
#pragma unroll
    for (int i = 0; i < DC * D; ++i) md[i] = 0.f;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    for (int64_t tt = (int64_t)blockIdx.x * kThreads + threadIdx.x; tt < a.rows * LPP; tt += nthr) {
        const int64_t p = tt / LPP;
        const int c0 = (int)(tt % LPP) * DC;
        const int64_t base = (p / a.Nq) * a.Ns;
        float g[DC];
        load_x8<D>(a.g, p, c0, g);
        for_edges(a.rel, a.idx, p, a.K, [&](float rx, float ry, float rz, int64_t j) {
            const int64_t row = base + j;
            float xj[DC], h1[D], pre1[D], a1[D], h2[DC], w[DC];
            load_x8<D>(a.x, row, c0, xj);
            edge_mlp<D>(sw, c0, rx, ry, rz, a.slope1, h1, pre1, a1, h2);
#pragma unroll
            for (int c = 0; c < DC; ++c) {
                const float dw = g[c] * xj[c];
                const float hh = (h2[c] - s_c[2][c0 + c]) * s_c[3][c0 + c];
                sm[c] += dw;
                sm[DC + c] = fmaf(dw, hh, sm[DC + c]);
                w[c] = fmaf(h2[c], s_c[0][c0 + c], s_c[1][c0 + c]) * g[c];
#pragma unroll
                for (int b = 0; b < D; ++b) md[c * D + b] = fmaf(dw, a1[b], md[c * D + b]);
            }
            if (a.dx) {
                red_add_v4(a.dx + row * D + c0, w[0], w[1], w[2], w[3]);
                red_add_v4(a.dx + row * D + c0 + 4, w[4], w[5], w[6], w[7]);
            }
        });
    }
    block_sums_to_slot<LPP>(sm, a.sums2, 2 * D, s_acc, [](int i, int blk) { return (i < DC ? 0 : D - DC) + blk * DC + i; });
    block_sums_to_slot<LPP>(md, a.mdw, D * D, s_acc, [](int i, int blk) { return (blk * DC + i / D) * D + i % D; });
}

// ---------------------------------------------------------------------------------------------------- backward, pass 2
struct Bwd2Args {
    const float* x; const float* rel; const int64_t* idx; const float* g;
    const float* W1; const float* W2; const float* sc1; const float* sh1; float slope1;
    const float* mu1; const float* is1;
    const float* sc2; const float* mu2; const float* is2; const float* k1b; const float* k2b;   // BatchNorm-2 backward constants (after its finalize)
    float* sums1;                           // [kStatSlots][2D]   Σdv1 | Σdv1·ĥ1
    float* s1;                              // [kStatSlots][3D]   Σ dv1[c]·r[a]
    int64_t rows, Ns, Nq; int K;
};

template <int D>
__global__ void __launch_bounds__(kThreads, D == 8 ? 3 : 2) bwd2_kernel(const Bwd2Args a) {
    constexpr int LPP = D / DC;
    __shared__ EdgeW<D> sw;
    __shared__ float4 s_w2t[D][D / 4];      // column b of W2: W2[0..D)[b]
    __shared__ float s_c[5][D];             // A = sc2, Bc, Cc (dh2 = A·dw + Bc + Cc·h2), mu1, is1
    __shared__ float s_acc[3 * D];
    stage_edge_w(sw, a.W1, a.W2, a.sc1, a.sh1);
    for (int i = threadIdx.x; i < D * D / 4; i += kThreads) {
        const int b = i / (D / 4), q = i % (D / 4);
        s_w2t[b][q] = make_float4(__ldg(a.W2 + (4 * q) * D + b), __ldg(a.W2 + (4 * q + 1) * D + b), __ldg(a.W2 + (4 * q + 2) * D + b),
                                  __ldg(a.W2 + (4 * q + 3) * D + b));
    }
    if (threadIdx.x < D) {
        const int c = threadIdx.x;
        const float sc = __ldg(a.sc2 + c), mu = __ldg(a.mu2 + c), is = __ldg(a.is2 + c), k1 = __ldg(a.k1b + c), k2 = __ldg(a.k2b + c);
        s_c[0][c] = sc; s_c[1][c] = -sc * k1 + sc * is * k2 * mu; s_c[2][c] = -sc * is * k2;
        s_c[3][c] = __ldg(a.mu1 + c); s_c[4][c] = __ldg(a.is1 + c);
    }
    __syncthreads();
    float sm[2 * DC], s1[3 * DC];
#pragma unroll
    for (int i = 0; i < 2 * DC; ++i) sm[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3 * DC; ++i) s1[i] = 0.f;
    const int64_t nthr = (int64_t)gridDim.x * kThreads;
    const int64_t total = a.rows * LPP;
    // all lanes of a warp run the same number of trips (the dh2 exchange below is a warp shuffle): clamp instead of exiting
    const int64_t trips = (total + nthr - 1) / nthr;
    for (int64_t it = 0; it < trips; ++it) {
        const int64_t t0 = (int64_t)blockIdx.x * kThreads + threadIdx.x + it * nthr;
        const bool live = t0 < total;
        const int64_t tt = live ? t0 : total - LPP + (t0 % LPP);          // a clamped thread keeps its channel block and adds nothing
        const int64_t p = tt / LPP;
        const int c0 = (int)(tt % LPP) * DC;
        const int64_t base = (p / a.Nq) * a.Ns;
        float g[DC];
        load_x8<D>(a.g, p, c0, g);
        for_edges(a.rel, a.idx, p, a.K, [&](float rx, float ry, float rz, int64_t j) {
            float xj[DC], h1[D], pre1[D], a1[D], h2[DC], dh2[D];
            load_x8<D>(a.x, base + j, c0, xj);
            edge_mlp<D>(sw, c0, rx, ry, rz, a.slope1, h1, pre1, a1, h2);
            float own[DC];
#pragma unroll
            for (int c = 0; c < DC; ++c) own[c] = fmaf(s_c[0][c0 + c], g[c] * xj[c], fmaf(s_c[2][c0 + c], h2[c], s_c[1][c0 + c]));
            if constexpr (LPP == 1) {
#pragma unroll
                for (int c = 0; c < DC; ++c) dh2[c] = own[c];
            } else {                                                     // the partner lane owns the other 8 channels of dh2
#pragma unroll
                for (int c = 0; c < DC; ++c) {
                    const float other = __shfl_xor_sync(0xffffffffu, own[c], 1);
                    dh2[c] = c0 == 0 ? own[c] : other;
                    dh2[DC + c] = c0 == 0 ? other : own[c];
                }
            }
#pragma unroll
            for (int b = 0; b < DC; ++b) {
                const int bb = c0 + b;
                float da = 0.f;
#pragma unroll
                for (int q = 0; q < D / 4; ++q) {
                    const float4 w = s_w2t[bb][q];
                    da = fmaf(w.x, dh2[4 * q], fmaf(w.y, dh2[4 * q + 1], fmaf(w.z, dh2[4 * q + 2], fmaf(w.w, dh2[4 * q + 3], da))));
                }
                const float pre = LPP == 1 ? pre1[b] : (c0 == 0 ? pre1[b] : pre1[DC + b]);
                const float hb = LPP == 1 ? h1[b] : (c0 == 0 ? h1[b] : h1[DC + b]);
                float dv = pre > 0.f ? da : da * a.slope1;
                if (!live) dv = 0.f;
                const float hh = (hb - s_c[3][bb]) * s_c[4][bb];
                sm[b] += dv;
                sm[DC + b] = fmaf(dv, hh, sm[DC + b]);
                s1[3 * b] = fmaf(dv, rx, s1[3 * b]); s1[3 * b + 1] = fmaf(dv, ry, s1[3 * b + 1]); s1[3 * b + 2] = fmaf(dv, rz, s1[3 * b + 2]);
            }
        });
    }
    block_sums_to_slot<LPP>(sm, a.sums1, 2 * D, s_acc, [](int i, int blk) { return (i < DC ? 0 : D - DC) + blk * DC + i; });
    block_sums_to_slot<LPP>(s1, a.s1, 3 * D, s_acc, [](int i, int blk) { return blk * 3 * DC + i; });
}

// ---------------------------------------------------------------------------------------------------- parameter gradients from the sums
struct ParamArgs {
    const float* mom; const float* asum; const float* mdw; const float* s1;
    const float* W1; const float* W2;
    const float* sc1; const float* mu1; const float* is1; const float* k1a; const float* k2a;    // BatchNorm 1 (after its backward finalize)
    const float* sc2; const float* mu2; const float* is2; const float* k1b; const float* k2b;    // BatchNorm 2
    float* dW1; float* dW2;                 // [D,3], [D,D]   +=
    int D;
};

__global__ void param_grads_kernel(const ParamArgs a) {
    extern __shared__ double sh_d[];        // mom [9] | asum [D + D·D] | mdw [D·D] | s1 [3D]
    const int D = a.D, tid = threadIdx.x, NA = D + D * D;
    double* m = sh_d;
    double* as = m + kMom;
    double* md = as + NA;
    double* s1 = md + D * D;
    for (int i = tid; i < kMom; i += blockDim.x) m[i] = fold_slots(a.mom, kMom, i);
    for (int i = tid; i < NA; i += blockDim.x) as[i] = fold_slots(a.asum, NA, i);
    for (int i = tid; i < D * D; i += blockDim.x) md[i] = fold_slots(a.mdw, D * D, i);
    for (int i = tid; i < 3 * D; i += blockDim.x) s1[i] = fold_slots(a.s1, 3 * D, i);
    __syncthreads();
    const bool upper = D == DC;             // d = 8 stores only the upper triangle of Σa1a1ᵀ
    for (int i = tid; i < D * D; i += blockDim.x) {            // dW2[c][b]
        const int c = i / D, b = i % D;
        double wsaa = 0.0;
        for (int q = 0; q < D; ++q) wsaa += (double)a.W2[c * D + q] * as[D + ((upper && q > b) ? b * D + q : q * D + b)];
        const double sc = a.sc2[c], mu = a.mu2[c], is = a.is2[c], k1 = a.k1b[c], k2 = a.k2b[c];
        a.dW2[i] += (float)(sc * (md[i] - k1 * as[b] - k2 * is * (wsaa - mu * as[b])));
    }
    for (int i = tid; i < 3 * D; i += blockDim.x) {            // dW1[c][x]
        const int c = i / 3, x = i % 3;
        double wsrr = 0.0;
        for (int q = 0; q < 3; ++q) wsrr += (double)a.W1[c * 3 + q] * sym3(m + 3, q, x);
        const double sc = a.sc1[c], mu = a.mu1[c], is = a.is1[c], k1 = a.k1a[c], k2 = a.k2a[c];
        a.dW1[i] += (float)(sc * (s1[i] - k1 * m[x] - k2 * is * (wsrr - mu * m[x])));
    }
}

template <int D>
inline unsigned point_grid(int64_t rows) {
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows * (D / DC), (int64_t)kThreads), (int64_t)kNumSMs * (D == 8 ? 3 : 2)));
}
inline bool width_ok(int D) { return D == 8 || D == 16; }

}  // namespace pcf
}  // namespace crf

using namespace crf;

extern "C" {

// 1 when the hidden width d is covered (8, 16)
int crfconv_pcf_supported(int D) { return pcf::width_ok(D) ? 1 : 0; }
// floats of the zero-initialised scratch: forward = BN1 statistics | BN2 statistics | activation sums (the moments are separate: 64·9);
// backward = Σdw sums | Σ dw a1ᵀ | Σdv1 sums | Σ dv1 rᵀ
int crfconv_pcf_fwd_scratch_floats(int D) { return kStatSlots * (2 * D + 2 * D + D + D * D); }
int crfconv_pcf_bwd_scratch_floats(int D) { return kStatSlots * (2 * D + D * D + 2 * D + 3 * D); }

// rel[e] = centre − support[idx[e]] and the moments Σr, Σrrᵀ (mom: [CRFCONV_STAT_SLOTS][9] zeroed floats)
int crfconv_pcf_relpos_moments(const float* support, const float* centres, const int64_t* idx, float* rel, float* mom, int64_t B, int64_t Ns,
                               int64_t Nq, int K, void* stream) {
    if (B <= 0 || Ns <= 0 || Nq <= 0 || K <= 0 || !support || !centres || !idx || !rel || !mom) return CRF_ERR_INVALID_ARG;
    const int64_t E = B * Nq * K;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(E, (int64_t)256), (int64_t)kNumSMs * 8);
    pcf::relpos_moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(support, centres, idx, rel, mom, E, Ns, Nq, K);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

// BN1 statistics of h1 = W1·r from the moments → stats1 [CRFCONV_STAT_SLOTS][2D] (slot 0; zeroed by the caller)
int crfconv_pcf_stats1(const float* mom, const float* W1, float* stats1, int D, void* stream) {
    if (!mom || !W1 || !stats1 || !pcf::width_ok(D)) return CRF_ERR_INVALID_ARG;
    pcf::stats1_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mom, W1, stats1, D);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pcf_fwd(const float* x, const float* rel, const int64_t* idx, const float* W1, const float* W2, const float* sc1, const float* sh1,
                    float slope1, float* P, float* Q, float* stats2, float* asum, int64_t B, int64_t Ns, int64_t Nq, int K, int D, void* stream) {
    if (B <= 0 || Ns <= 0 || Nq <= 0 || K <= 0 || !x || !rel || !idx || !W1 || !W2 || !sc1 || !sh1 || !P || !Q || !stats2 || !asum)
        return CRF_ERR_INVALID_ARG;
    if (!pcf::width_ok(D)) return CRF_ERR_UNSUPPORTED;
    pcf::FwdArgs a{x, rel, idx, W1, W2, sc1, sh1, slope1, P, Q, stats2, asum, B * Nq, Ns, Nq, K};
    if (D == 8) pcf::fwd_kernel<8><<<pcf::point_grid<8>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    else pcf::fwd_kernel<16><<<pcf::point_grid<16>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pcf_out(const float* P, const float* Q, const float* sc2, const float* sh2, float* out, int64_t rows, int D, void* stream) {
    if (rows <= 0 || !P || !Q || !sc2 || !sh2 || !out || !pcf::width_ok(D)) return CRF_ERR_INVALID_ARG;
    const int64_t total4 = rows * (D / 4);
    pcf::out_kernel<<<(unsigned)std::min<int64_t>(ceil_div(total4, (int64_t)256), (int64_t)kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(P, Q, sc2, sh2, out, total4, D / 4);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pcf_bwd1(const float* x, const float* rel, const int64_t* idx, const float* g, const float* W1, const float* W2, const float* sc1,
                     const float* sh1, float slope1, const float* sc2, const float* sh2, const float* mu2, const float* is2, float* dx,
                     float* sums2, float* mdw, int64_t B, int64_t Ns, int64_t Nq, int K, int D, void* stream) {
    if (B <= 0 || Ns <= 0 || Nq <= 0 || K <= 0 || !x || !rel || !idx || !g || !W1 || !W2 || !sc1 || !sh1 || !sc2 || !sh2 || !mu2 || !is2 || !sums2 || !mdw)
        return CRF_ERR_INVALID_ARG;
    if (!pcf::width_ok(D)) return CRF_ERR_UNSUPPORTED;
    pcf::Bwd1Args a{x, rel, idx, g, W1, W2, sc1, sh1, slope1, sc2, sh2, mu2, is2, dx, sums2, mdw, B * Nq, Ns, Nq, K};
    if (D == 8) pcf::bwd1_kernel<8><<<pcf::point_grid<8>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    else pcf::bwd1_kernel<16><<<pcf::point_grid<16>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

int crfconv_pcf_bwd2(const float* x, const float* rel, const int64_t* idx, const float* g, const float* W1, const float* W2, const float* sc1,
                     const float* sh1, float slope1, const float* mu1, const float* is1, const float* sc2, const float* mu2, const float* is2,
                     const float* k1b, const float* k2b, float* sums1, float* s1, int64_t B, int64_t Ns, int64_t Nq, int K, int D, void* stream) {
    if (B <= 0 || Ns <= 0 || Nq <= 0 || K <= 0 || !x || !rel || !idx || !g || !W1 || !W2 || !sc1 || !sh1 || !mu1 || !is1 || !sc2 || !mu2 || !is2 || !k1b ||
        !k2b || !sums1 || !s1)
        return CRF_ERR_INVALID_ARG;
    if (!pcf::width_ok(D)) return CRF_ERR_UNSUPPORTED;
    pcf::Bwd2Args a{x, rel, idx, g, W1, W2, sc1, sh1, slope1, mu1, is1, sc2, mu2, is2, k1b, k2b, sums1, s1, B * Nq, Ns, Nq, K};
    if (D == 8) pcf::bwd2_kernel<8><<<pcf::point_grid<8>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    else pcf::bwd2_kernel<16><<<pcf::point_grid<16>(a.rows), pcf::kThreads, 0, (cudaStream_t)stream>>>(a);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

// dW1 [D,3] and dW2 [D,D] (+=) from the sums of the forward and the two backward passes and the finalized BatchNorm-backward constants
int crfconv_pcf_param_grads(const float* mom, const float* asum, const float* mdw, const float* s1, const float* W1, const float* W2,
                            const float* sc1, const float* mu1, const float* is1, const float* k1a, const float* k2a, const float* sc2,
                            const float* mu2, const float* is2, const float* k1b, const float* k2b, float* dW1, float* dW2, int D, void* stream) {
    if (!mom || !asum || !mdw || !s1 || !W1 || !W2 || !sc1 || !mu1 || !is1 || !k1a || !k2a || !sc2 || !mu2 || !is2 || !k1b || !k2b || !dW1 || !dW2)
        return CRF_ERR_INVALID_ARG;
    if (!pcf::width_ok(D)) return CRF_ERR_UNSUPPORTED;
    pcf::ParamArgs a{mom, asum, mdw, s1, W1, W2, sc1, mu1, is1, k1a, k2a, sc2, mu2, is2, k1b, k2b, dW1, dW2, D};
    const size_t smem = (size_t)(pcf::kMom + D + D * D + D * D + 3 * D) * sizeof(double);
    pcf::param_grads_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(a);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

}  // extern "C"
