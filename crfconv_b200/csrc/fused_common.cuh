// Building blocks of the layer-specialised ("fused") CRF kernels in crf_fused.cu and of the grid-wide BatchNorm
// bookkeeping that the GEMM kernels run in their own tail instead of in separate launches:
//   * 3xTF32 mma.sync helpers with a truncating (hi, lo) split — fp32-grade products (≈2^-21) for the narrow contractions;
//   * "last CTA finishes the reduction": every CTA publishes its partial sums, takes a ticket, and the CTA that draws the
//     last ticket folds all partials in a FIXED order and runs the BatchNorm finalize (forward: scale/shift/mean/invstd and
//     the running statistics of nn.BatchNorm1d; backward: k1/k2 and dγ/dβ).  This removes the 12 one-CTA finalize launches
//     of a CRF layer step (≈4.5 us each on the critical path) without giving up run-to-run reproducible statistics.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"
#include "mma.cuh"

namespace crf {
namespace cl {

// ------------------------------------------------------------------------------------------------ BatchNorm finalize blocks
struct FwdFin {                 // forward, training mode (batch statistics).  part == nullptr ⇒ disabled
    float* part;                // [nparts][2C] partial Σ | Σ²
    unsigned int* counter;      // kTicketInts zeroed uint32 (arrive_is_last); left zero
    const float* gamma; const float* beta;   // may be null (1 / 0)
    float* running_mean; float* running_var; // may be null
    float eps, momentum;
    double count;               // rows
    float* scale; float* shift; float* mean; float* invstd;   // outputs [C]
};

struct BwdFin {                 // backward: s1 = Σ dV, s2 = Σ dV·Ĥ.  part == nullptr ⇒ disabled
    float* part;                // [nparts][2C]
    unsigned int* counter;
    double count;
    float* k1; float* k2;       // outputs [C]: s1/count, s2/count
    float* dgamma; float* dbeta;   // += s2, += s1 (may be null)
};

// Every thread of the CTA calls this after its partial results are written to global memory.  Returns true in exactly one CTA
// of the grid: the one that arrives last, at which point all partials of all CTAs are visible to it.
// Tickets are two-level: same-address atomics serialise at ≈40 ns each in L2, and the CTAs of a persistent grid all finish within
// a microsecond of each other — a single counter turned every kernel's tail into (#CTAs × 40 ns) ≈ 6-50 us.  CTAs draw a ticket
// from the counter of their group of 16; the last of each group draws one from the second-level counter: ≤ 16 + 32 serialised
// atomics.  `counters` = kTicketInts zeroed uint32 (left zero); gridDim.x ≤ 16·(kTicketInts − 1).
constexpr int kTicketInts = 40;
constexpr int kMaxTicketGrid = 16 * (kTicketInts - 1);
__device__ __forceinline__ bool arrive_is_last(unsigned int* counters) {
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned grp = blockIdx.x >> 4, ngrp = (gridDim.x + 15) >> 4;
        const unsigned in_grp = min(16u, gridDim.x - (grp << 4));
        unsigned last = 0;
        if (atomicAdd(counters + 1 + grp, 1u) == in_grp - 1) {
            counters[1 + grp] = 0u;                              // nobody else touches this counter again in this launch
            __threadfence();
            if (atomicAdd(counters, 1u) == ngrp - 1) { counters[0] = 0u; last = 1u; }
        }
        s_last = last;
    }
    __syncthreads();
    const bool last = s_last != 0u;
    if (last) __threadfence();
    return last;
}

// Sums V values per partial row over `nparts` rows in a fixed order.  Thread (v = tid % V, sub = tid / V), tid < NT, adds rows
// sub, sub + NT/V, ... in double; the NT/V sub-sums are combined through `red` (NT doubles of shared memory).  On return
// red[0..V) holds the totals.  Called by ALL threads of the (one) finishing CTA; blockDim.x >= NT, NT % V == 0.
template <int V, int NT>
__device__ __forceinline__ void sum_parts(const float* part, int nparts, double* red) {
    constexpr int NSUB = NT / V;
    static_assert(NT % V == 0, "NT must be a multiple of V");
    const int tid = threadIdx.x;
    if (tid < NT) {
        const int v = tid % V, sub = tid / V;
        double acc = 0.0;
        for (int p0 = sub; p0 < nparts; p0 += NSUB * 8) {
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int p = p0 + u * NSUB;
                x[u] = p < nparts ? __ldcg(part + (size_t)p * V + v) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += (double)x[u];
        }
        red[tid] = acc;
    }
    __syncthreads();
    if (tid < V) {
        double t = 0.0;
#pragma unroll
        for (int s = 0; s < NSUB; ++s) t += red[s * V + tid];
        red[tid] = t;
    }
    __syncthreads();
}

// nn.BatchNorm1d training-mode bookkeeping for channel c from the totals s1 = Σx, s2 = Σx² (same arithmetic as
// lin::bn_finalize_fwd_kernel, linear.cu).
__device__ __forceinline__ void bn_fwd_finalize(const FwdFin& f, double s1, double s2, int c) {
    const double mu = s1 / f.count;
    double var = s2 / f.count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mu;
    const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
    if (f.running_mean) {
        const double unb = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
        f.running_mean[c] = (1.0f - f.momentum) * f.running_mean[c] + f.momentum * (float)mu;
        f.running_var[c] = (1.0f - f.momentum) * f.running_var[c] + f.momentum * (float)unb;
    }
    const float gm = f.gamma ? f.gamma[c] : 1.0f, bt = f.beta ? f.beta[c] : 0.0f;
    const float sc = gm * invstd;
    f.scale[c] = sc;
    f.shift[c] = bt - mean * sc;
    if (f.mean) f.mean[c] = mean;
    if (f.invstd) f.invstd[c] = invstd;
}

__device__ __forceinline__ void bn_bwd_finalize(const BwdFin& f, double s1, double s2, int c) {
    f.k1[c] = (float)(s1 / f.count);
    f.k2[c] = (float)(s2 / f.count);
    if (f.dgamma) f.dgamma[c] += (float)s2;
    if (f.dbeta) f.dbeta[c] += (float)s1;
}

// Tail of a kernel whose CTAs have added into `nparts` zero-initialised slot rows part[slot][2C] (atomics): the last CTA to arrive
// folds the slots and finalizes.  Must be reached by every thread of every CTA.
template <int C, int NT>
__device__ __forceinline__ void fwd_fin_tail(const FwdFin& f, int nparts, double* red) {
    if (!arrive_is_last(f.counter)) return;
    sum_parts<2 * C, NT>(f.part, nparts, red);
    if (threadIdx.x < C) bn_fwd_finalize(f, red[threadIdx.x], red[C + threadIdx.x], threadIdx.x);
}
template <int C, int NT>
__device__ __forceinline__ void bwd_fin_tail(const BwdFin& f, int nparts, double* red) {
    if (!arrive_is_last(f.counter)) return;
    sum_parts<2 * C, NT>(f.part, nparts, red);
    if (threadIdx.x < C) bn_bwd_finalize(f, red[threadIdx.x], red[C + threadIdx.x], threadIdx.x);
}

// Deterministic grid reduction of per-CTA rows (V floats each, written with plain stores to rows[blockIdx.x][V]) with the two-level
// arrival of arrive_is_last: the last CTA of each group of 16 folds its group's rows — fixed order, double precision — into one group
// row, and the last group folds the ≤ kTicketInts − 1 group rows.  Compared with one CTA folding gridDim.x rows this cuts the serial
// tail of a kernel from ≈(gridDim.x / 8) dependent L2 round trips to 2-3.  On return true (exactly one CTA) red[0..V) holds the totals.
// rows: gridDim.x·V floats;  grows: (kTicketInts − 1)·V doubles;  red: NT doubles of shared memory;  NT % V == 0, blockDim.x ≥ NT.
template <int V, int NT>
__device__ __forceinline__ bool grid_reduce_rows(const float* rows, double* grows, unsigned int* counters, double* red) {
    __shared__ unsigned int s_flag;
    constexpr int NSUB = NT / V;
    static_assert(NT % V == 0, "NT must be a multiple of V");
    const int tid = threadIdx.x;
    const unsigned grp = blockIdx.x >> 4, ngrp = (gridDim.x + 15) >> 4;
    const unsigned in_grp = min(16u, gridDim.x - (grp << 4));
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const bool glast = atomicAdd(counters + 1 + grp, 1u) == in_grp - 1;
        if (glast) counters[1 + grp] = 0u;
        s_flag = glast ? 1u : 0u;
    }
    __syncthreads();
    if (s_flag == 0u) return false;
    __threadfence();
    // ---- fold this group's rows (≤ 16) in CTA order
    if (tid < NT) {
        const int v = tid % V, sub = tid / V;
        float x[(16 + NSUB - 1) / NSUB];
#pragma unroll
        for (int u = 0; u < (16 + NSUB - 1) / NSUB; ++u) {
            const unsigned r = sub + u * NSUB;
            x[u] = r < in_grp ? __ldcg(rows + ((size_t)(grp << 4) + r) * V + v) : 0.f;
        }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < (16 + NSUB - 1) / NSUB; ++u) acc += (double)x[u];
        red[tid] = acc;
    }
    __syncthreads();
    if (tid < V) {
        double t = 0.0;
#pragma unroll
        for (int s = 0; s < NSUB; ++s) t += red[s * V + tid];
        grows[(size_t)grp * V + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const bool last = atomicAdd(counters, 1u) == ngrp - 1;
        if (last) counters[0] = 0u;
        s_flag = last ? 1u : 0u;
    }
    __syncthreads();
    if (s_flag == 0u) return false;
    __threadfence();
    // ---- fold the group rows in group order
    if (tid < NT) {
        const int v = tid % V, sub = tid / V;
        double acc = 0.0;
        for (unsigned r0 = sub; r0 < ngrp; r0 += NSUB * 4) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned r = r0 + u * NSUB;
                x[u] = r < ngrp ? __ldcg(grows + (size_t)r * V + v) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += x[u];
        }
        red[tid] = acc;
    }
    __syncthreads();
    if (tid < V) {
        double t = 0.0;
#pragma unroll
        for (int s = 0; s < NSUB; ++s) t += red[s * V + tid];
        red[tid] = t;
    }
    __syncthreads();
    return true;
}
// layout of a `part` scratch used with grid_reduce_rows: [kMaxTicketGrid][V] floats, then [kTicketInts − 1][V] doubles (V ≤ 32)
__device__ __forceinline__ double* part_group_rows(float* part) { return reinterpret_cast<double*>(part + (size_t)kMaxTicketGrid * 32); }
constexpr int kPartFloats = kMaxTicketGrid * 32 + (kTicketInts - 1) * 32 * 2;

// ------------------------------------------------------------------------------------------------ cp.async (warp-private tile rings)
// 16-byte asynchronous global→shared copy; !pred ⇒ the destination is zero-filled and the source is not read
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------ 3xTF32 mma.sync helpers
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.0f ? v : v * slope; }

// hi = the 19 leading bits the tensor core reads (sign, exponent, 10 mantissa bits), lo = x − hi exactly
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
struct FragA { uint32_t hi[4], lo[4]; };
__device__ __forceinline__ void make_a(FragA& f, float a0, float a1, float a2, float a3) {
    split(a0, f.hi[0], f.lo[0]); split(a1, f.hi[1], f.lo[1]); split(a2, f.hi[2], f.lo[2]); split(a3, f.hi[3], f.lo[3]);
}
struct FragB { uint32_t hi[2], lo[2]; };
__device__ __forceinline__ void make_b(FragB& f, float b0, float b1) { split(b0, f.hi[0], f.lo[0]); split(b1, f.hi[1], f.lo[1]); }
// D += A·B to fp32 accuracy: small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const FragA& a, const FragB& b) {
    mma_tf32(d, a.lo, b.hi);
    mma_tf32(d, a.hi, b.lo);
    mma_tf32(d, a.hi, b.hi);
}
// B fragments kept pre-split in shared memory as (hi0, hi1) / (lo0, lo1) float2 pairs, one pair per lane: conflict-free LDS.64
__device__ __forceinline__ void mma3(float (&d)[4], const FragA& a, float2 bh, float2 bl) {
    FragB b;
    b.hi[0] = __float_as_uint(bh.x); b.hi[1] = __float_as_uint(bh.y);
    b.lo[0] = __float_as_uint(bl.x); b.lo[1] = __float_as_uint(bl.y);
    mma3(d, a, b);
}
__device__ __forceinline__ void store_split(float2* Bh, float2* Bl, int i, float w0, float w1) {
    uint32_t h0, l0, h1, l1;
    split(w0, h0, l0);
    split(w1, h1, l1);
    Bh[i] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
    Bl[i] = make_float2(__uint_as_float(l0), __uint_as_float(l1));
}

// ---- a 16×16 matrix applied to the 8 point rows a warp holds in the "4 lanes × float4" layout ----------------------------------
// Lane (g, t) = (lane >> 2, lane & 3) holds channels 4t..4t+3 of point g.  That IS an m16n8k8 A fragment (rows 8..15 zero) once
// the k index is permuted (k-step s: k = t ↔ channel 4t+2s, k = t+4 ↔ channel 4t+2s+1), and with the output columns permuted by
// phys_col the two accumulator blocks come back as channels 4t..4t+3 of point g: row-vector × matrix products chain through
// registers with 12 MMAs each, no shuffle and no operand broadcast from shared memory.  (The shuffle + LDS.128 form costs
// 16 + 4·16 = 80 L1 data-pipe wavefronts per product and warp — the mean-field kernels are bound by that pipe — this one 16.)
// Matrix fragments live pre-split in shared memory: frag[(s·2 + nb)·32 + lane] = { hi(b0), hi(b1), lo(b0), lo(b1) }.
template <typename GetW>
__device__ __forceinline__ void stage_mat16(float4* frag, GetW w, int tid, int nthreads) {   // w(k, n) = Mat[k][n] of  out = x·Mat
    for (int i = tid; i < 128; i += nthreads) {
        const int s = i >> 6, nb = (i >> 5) & 1, gg = (i & 31) >> 2, tt = i & 3;
        const int kin = 4 * tt + 2 * s, nout = 4 * (gg >> 1) + 2 * nb + (gg & 1);
        uint32_t h0, l0, h1, l1;
        split(w(kin, nout), h0, l0);
        split(w(kin + 1, nout), h1, l1);
        frag[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
}
__device__ __forceinline__ float4 rows8_mat16(float4 x, const float4* frag, int lane) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    FragA f;
    f.hi[1] = f.lo[1] = f.hi[3] = f.lo[3] = 0u;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        split(s ? x.z : x.x, f.hi[0], f.lo[0]);
        split(s ? x.w : x.y, f.hi[2], f.lo[2]);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const float4 b = frag[(s * 2 + nb) * 32 + lane];
            mma3(acc[nb], f, make_float2(b.x, b.y), make_float2(b.z, b.w));
        }
    }
    return make_float4(acc[0][0], acc[0][1], acc[1][0], acc[1][1]);
}

// Output-column permutation that turns the two 8-wide accumulator blocks 2q, 2q+1 of a thread into ONE float4 of 4 consecutive
// physical columns 16q + 4t .. 16q + 4t + 3 (so stores / read-modify-writes are 128-bit and line up with row-major float4 loads):
//   MMA column n of block nb  ↔  physical column 16·(nb>>1) + 4·(n>>1) + 2·(nb&1) + (n&1)
__host__ __device__ __forceinline__ int phys_col(int nb, int n) { return 16 * (nb >> 1) + 4 * (n >> 1) + 2 * (nb & 1) + (n & 1); }

}  // namespace cl
}  // namespace crf
