// Exact batched kNN on sm_100a — the B200-native replacement of the reference's nanoflann wrapper
// (/root/reference/utils/nearest_neighbors/knn_.cxx:22-135, contract in SURVEY.md Appendix A.1).
//
// Result contract (bit-exact): for every query the K support points with the smallest
//     d = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)),   dx = fl(q.x - p.x) ...        (nanoflann.hpp:343-346, no FMA)
// in ascending (d, index) order.  The kd-tree of the reference is NOT re-implemented: each cloud is binned into a
// uniform grid (counting sort into cell-contiguous float4 records), and one warp per query scans growing shells of
// cells keeping the running top-32 in registers (one slot per lane, shuffle-insert).  The search stops when the
// K-th distance is strictly below a lower bound on the *computed* distance of every unvisited point; the bound is
// derived from per-cell edge tables that are valid for the exact f32 binning function (monotonicity argument in
// DESIGN.md §kNN), so the result is independent of the grid and equals a brute-force scan.
#include "../../include/crfconv_b200.h"
#include "common.cuh"

namespace crf {
namespace knn {

constexpr int kMaxDim = 1024;            // max grid cells per axis
constexpr int kEdgeStride = 3 * kMaxDim; // floats per edge table per cloud

struct Grid {
    float lo[3];
    float inv[3];
    int n[3];
    int ncells;
};

__host__ __device__ inline int64_t cell_cap(int64_t N) { return 2 * N + 4096; }

// The binning function.  Monotone non-decreasing in x (every step is a monotone rounding), which is all the
// correctness argument needs.
__device__ __forceinline__ int cell_coord(float x, float lo, float inv, int n) {
    float t = floorf(__fmul_rn(__fsub_rn(x, lo), inv));
    // clamp in float first: t may be huge / negative for queries outside the support bbox
    t = fminf(fmaxf(t, 0.0f), (float)(n - 1));
    return (int)t;
}

// nanoflann.hpp:343-346 arithmetic, contraction disabled explicitly.
__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------- build
__global__ void init_kernel(unsigned* bbox, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 6) bbox[i] = (i % 6 < 3) ? 0xffffffffu : 0u;   // [min xyz | max xyz] in ordered-uint space
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pts, unsigned* bbox, int N) {
    const int b = blockIdx.y;
    const float* p = pts + (size_t)b * N * 3;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __ldg(p + 3 * (size_t)i + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(bbox + b * 6 + c, f2ord(mn[c]));
            atomicMax(bbox + b * 6 + 3 + c, f2ord(mx[c]));
        }
    }
}

// One block per cloud: thread 0 chooses the grid, then the block fills the edge tables.
//   lo_edge[a][c] (c>=1): a float L with cell(L) <= c-1  ⇒ every point binned at >= c has coordinate > L
//   hi_edge[a][c] (c<=n-2): a float H with cell(H) >= c+1 ⇒ every point binned at <= c has coordinate < H
__global__ void __launch_bounds__(256) setup_kernel(const unsigned* __restrict__ bbox, Grid* grids, float* edges,
                                                    int N, float occ) {
    const int b = blockIdx.x;
    __shared__ Grid g;
    __shared__ float cell_sz[3];
    if (threadIdx.x == 0) {
        double ext[3], vol = 1.0;
        int nd = 0;
        for (int c = 0; c < 3; ++c) {
            float lo = ord2f(bbox[b * 6 + c]), hi = ord2f(bbox[b * 6 + 3 + c]);
            g.lo[c] = lo;
            ext[c] = (double)hi - (double)lo;
            if (!(ext[c] > 0.0)) ext[c] = 0.0;
            if (ext[c] > 0.0) { vol *= ext[c]; ++nd; }
        }
        double target = fmax(1.0, (double)N / (double)occ);
        double cell = nd ? pow(vol / target, 1.0 / nd) : 1.0;
        if (!(cell > 0.0) || !isfinite(cell)) cell = 1.0;
        const double cap = (double)cell_cap(N);
        for (int it = 0; it < 200; ++it) {
            double tot = 1.0;
            for (int c = 0; c < 3; ++c) {
                double nn = ext[c] > 0.0 ? floor(ext[c] / cell) + 1.0 : 1.0;
                if (nn > kMaxDim) nn = kMaxDim;
                g.n[c] = (int)nn;
                tot *= nn;
            }
            if (tot <= cap) break;
            cell *= 1.25;
        }
        for (int c = 0; c < 3; ++c) {
            double cs = cell;
            if (ext[c] > 0.0 && floor(ext[c] / cs) + 1.0 > kMaxDim) cs = ext[c] / kMaxDim * 1.0001;   // capped axis
            float inv = (float)(1.0 / cs);
            if (!isfinite(inv) || !(inv > 0.0f)) inv = 1.0f;
            g.inv[c] = inv;
            cell_sz[c] = (float)cs;
        }
        g.ncells = g.n[0] * g.n[1] * g.n[2];
        grids[b] = g;
    }
    __syncthreads();
    float* lo_edge = edges + (size_t)b * 2 * kEdgeStride;
    float* hi_edge = lo_edge + kEdgeStride;
    for (int t = threadIdx.x; t < 3 * kMaxDim; t += blockDim.x) {
        const int a = t / kMaxDim, c = t % kMaxDim;
        const int n = g.n[a];
        const float lo = g.lo[a], inv = g.inv[a], cs = cell_sz[a];
        float L = -INFINITY, H = INFINITY;
        if (c < n) {
            if (c >= 1) {
                L = __fadd_rn(lo, __fmul_rn((float)c, cs));
                float step = fmaxf(fabsf(L) * 1.2e-7f, cs * 1e-6f);
                int guard = 0;
                while (cell_coord(L, lo, inv, n) > c - 1 && guard++ < 64) { L = __fsub_rn(L, step); step *= 2.0f; }
                if (cell_coord(L, lo, inv, n) > c - 1) L = -INFINITY;
            }
            if (c <= n - 2) {
                H = __fadd_rn(lo, __fmul_rn((float)(c + 1), cs));
                float step = fmaxf(fabsf(H) * 1.2e-7f, cs * 1e-6f);
                int guard = 0;
                while (cell_coord(H, lo, inv, n) < c + 1 && guard++ < 64) { H = __fadd_rn(H, step); step *= 2.0f; }
                if (cell_coord(H, lo, inv, n) < c + 1) H = INFINITY;
            }
        }
        lo_edge[t] = L;
        hi_edge[t] = H;
    }
}

__global__ void __launch_bounds__(256) count_kernel(const float* __restrict__ pts, const Grid* __restrict__ grids,
                                                    int* __restrict__ cell_of, int* cell_count, int N, int cap) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Grid& g = grids[b];
    const float* p = pts + ((size_t)b * N + i) * 3;
    const int cx = cell_coord(__ldg(p), g.lo[0], g.inv[0], g.n[0]);
    const int cy = cell_coord(__ldg(p + 1), g.lo[1], g.inv[1], g.n[1]);
    const int cz = cell_coord(__ldg(p + 2), g.lo[2], g.inv[2], g.n[2]);
    const int cell = (cz * g.n[1] + cy) * g.n[0] + cx;
    cell_of[(size_t)b * N + i] = cell;
    atomicAdd(cell_count + (size_t)b * (cap + 1) + cell, 1);
}

// Exclusive scan of cell_count[b][0..ncells) in place (→ cell_start), one block per cloud; entry [ncells] = N.
__global__ void __launch_bounds__(1024) scan_kernel(int* cell_count, const Grid* __restrict__ grids, int cap) {
    const int b = blockIdx.x;
    int* a = cell_count + (size_t)b * (cap + 1);
    const int n = grids[b].ncells + 1;   // include the sentinel slot (holds 0 before the scan)
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? a[i] : 0;
        int x = v;
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
            warp_tot[lane] = t;   // inclusive
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (w ? warp_tot[w - 1] : 0) + x - v;
        if (i < n) a[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(const float* __restrict__ pts, const int* __restrict__ cell_of,
                                                      const int* __restrict__ cell_start, int* cursor,
                                                      float4* __restrict__ sorted, int N, int cap) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int cell = cell_of[(size_t)b * N + i];
    const int pos = cell_start[(size_t)b * (cap + 1) + cell] + atomicAdd(cursor + (size_t)b * cap + cell, 1);
    const float* p = pts + ((size_t)b * N + i) * 3;
    sorted[(size_t)b * N + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float(i));
}

// ---------------------------------------------------------------------------------------------- query
struct TopK {   // one slot per lane, ascending (d, i) by lane; slot K-1 is the current worst accepted
    float d;
    int i;
    float kth_d;
    int kth_i;
};

// (ld, li): lower threshold of a multi-pass search — only candidates strictly greater than it in (d, index) order may enter
// (first pass: ld = -1, nothing is excluded since distances are >= 0).
template <bool GENERAL>
__device__ __forceinline__ void scan_range(int s, int e, float qx, float qy, float qz,
                                           const float4* __restrict__ sorted, TopK& t, int K, int lane, float ld = -1.0f, int li = -1,
                                           float r2 = -1.0f) {
    for (int base = s; base < e; base += 32) {
        const int p = base + lane;
        float d = INFINITY;
        int idx = 0x7fffffff;
        if (p < e) {
            const float4 v = __ldg(sorted + p);
            d = sqdist(qx, qy, qz, v.x, v.y, v.z);
            idx = __float_as_int(v.w);
            // radius mode (r2 >= 0): every support point within the radius gets the same key 0, so the (d, index) order below keeps
            // the K SMALLEST INDICES among them — torch_cluster's "first max_num_neighbors in index order"
            if (GENERAL && r2 >= 0.0f) { if (d <= r2) d = 0.0f; else { d = INFINITY; idx = 0x7fffffff; } }
        }
        bool pass = (d < t.kth_d) || (d == t.kth_d && idx < t.kth_i);
        if (GENERAL) pass = pass && ((d > ld) || (d == ld && idx > li));
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float dc = __shfl_sync(0xffffffffu, d, src);
            const int ic = __shfl_sync(0xffffffffu, idx, src);
            if (!((dc < t.kth_d) || (dc == t.kth_d && ic < t.kth_i))) continue;   // worst may have tightened
            const bool less = (t.d < dc) || (t.d == dc && t.i < ic);
            const int pos = __popc(__ballot_sync(0xffffffffu, less));             // slots are sorted ⇒ prefix mask
            const float ud = __shfl_up_sync(0xffffffffu, t.d, 1);
            const int ui = __shfl_up_sync(0xffffffffu, t.i, 1);
            if (lane == pos) { t.d = dc; t.i = ic; }
            else if (lane > pos) { t.d = ud; t.i = ui; }
            t.kth_d = __shfl_sync(0xffffffffu, t.d, K - 1);
            t.kth_i = __shfl_sync(0xffffffffu, t.i, K - 1);
        }
    }
}

// SELF = queries are the support points themselves (same buffer): queries are then taken in CELL order from the sorted records,
// so the warps of a CTA search neighbouring cells and share their point ranges through L1; results go to the original row.
// GENERAL = false: the plain K <= 32 search (one pass, no radius) — the hot path of the multiscale builder, kept free of the
// threshold / radius logic; GENERAL = true: passes of 32 for K > 32 and the radius mode.
template <bool SELF, bool GENERAL>
__global__ void __launch_bounds__(256, 4) query_kernel(const float4* __restrict__ sorted_all,
                                                    const int* __restrict__ cell_start_all,
                                                    const Grid* __restrict__ grids, const float* __restrict__ edges,
                                                    const float* __restrict__ queries, int64_t* __restrict__ out,
                                                    int N, int Q, int K, int cap, float r2) {
    const int b = blockIdx.y;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int lane = lane_id();
    const Grid g = grids[b];
    const float4* sorted = sorted_all + (size_t)b * N;
    const int* cell_start = cell_start_all + (size_t)b * (cap + 1);
    const float* lo_edge = edges + (size_t)b * 2 * kEdgeStride;
    const float* hi_edge = lo_edge + kEdgeStride;

    float qx, qy, qz;
    int out_row = q;
    if (SELF) {
        const float4 me = __ldg(sorted + q);
        qx = me.x; qy = me.y; qz = me.z;
        out_row = __float_as_int(me.w);
    } else {
        const float* qp = queries + ((size_t)b * Q + q) * 3;
        qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
    }
    const int cx = cell_coord(qx, g.lo[0], g.inv[0], g.n[0]);
    const int cy = cell_coord(qy, g.lo[1], g.inv[1], g.n[1]);
    const int cz = cell_coord(qz, g.lo[2], g.inv[2], g.n[2]);
    const int nx = g.n[0], ny = g.n[1], nz = g.n[2];

    // K > 32: passes of 32 — pass p finds ranks 32p .. 32p+31 as the smallest candidates strictly above the last one of pass p-1
    // (each pass restarts from the query's own cell; the stopping rule below holds for the filtered candidate set as well)
    float ld = -1.0f;
    int li = -1;
    for (int k0 = 0; k0 < (GENERAL ? K : 1); k0 += 32) {
    const int Kp = GENERAL ? min(32, K - k0) : K;
    TopK t;
    t.d = INFINITY; t.i = 0x7fffffff; t.kth_d = INFINITY; t.kth_i = 0x7fffffff;

    for (int r = 0;; ++r) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, nx - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, ny - 1);
        const int z0 = max(cz - r, 0), z1 = min(cz + r, nz - 1);
        const int nyr = y1 - y0 + 1, nrows = nyr * (z1 - z0 + 1);
        const bool xa_ok = (cx - r >= 0), xb_ok = (cx + r <= nx - 1) && r > 0;
        for (int rb = 0; rb < nrows; rb += 32) {
            // each lane fetches the point ranges of one (y,z) row of the shell
            int sA = 0, eA = 0, sB = 0, eB = 0;
            const int j = rb + lane;
            if (j < nrows) {
                const int y = y0 + j % nyr, z = z0 + j / nyr;
                const int row = (z * ny + y) * nx;
                const bool face = (r == 0) || y == cy - r || y == cy + r || z == cz - r || z == cz + r;
                if (face) {
                    sA = __ldg(cell_start + row + x0);
                    eA = __ldg(cell_start + row + x1 + 1);
                } else {
                    if (xa_ok) { sA = __ldg(cell_start + row + cx - r); eA = __ldg(cell_start + row + cx - r + 1); }
                    if (xb_ok) { sB = __ldg(cell_start + row + cx + r); eB = __ldg(cell_start + row + cx + r + 1); }
                }
            }
            const int nb = min(32, nrows - rb);
            for (int l = 0; l < nb; ++l) {
                const int a0 = __shfl_sync(0xffffffffu, sA, l), a1 = __shfl_sync(0xffffffffu, eA, l);
                const int b0 = __shfl_sync(0xffffffffu, sB, l), b1 = __shfl_sync(0xffffffffu, eB, l);
                if (a1 > a0) scan_range<GENERAL>(a0, a1, qx, qy, qz, sorted, t, Kp, lane, ld, li, r2);
                if (b1 > b0) scan_range<GENERAL>(b0, b1, qx, qy, qz, sorted, t, Kp, lane, ld, li, r2);
            }
        }
        // lower bound on the computed distance of every point outside the visited box
        float bound = INFINITY;
        if (cx + r + 1 <= nx - 1) bound = fminf(bound, fmaxf(0.0f, __fsub_rn(__ldg(lo_edge + cx + r + 1), qx)));
        if (cx - r - 1 >= 0)      bound = fminf(bound, fmaxf(0.0f, __fsub_rn(qx, __ldg(hi_edge + cx - r - 1))));
        if (cy + r + 1 <= ny - 1) bound = fminf(bound, fmaxf(0.0f, __fsub_rn(__ldg(lo_edge + kMaxDim + cy + r + 1), qy)));
        if (cy - r - 1 >= 0)      bound = fminf(bound, fmaxf(0.0f, __fsub_rn(qy, __ldg(hi_edge + kMaxDim + cy - r - 1))));
        if (cz + r + 1 <= nz - 1) bound = fminf(bound, fmaxf(0.0f, __fsub_rn(__ldg(lo_edge + 2 * kMaxDim + cz + r + 1), qz)));
        if (cz - r - 1 >= 0)      bound = fminf(bound, fmaxf(0.0f, __fsub_rn(qz, __ldg(hi_edge + 2 * kMaxDim + cz - r - 1))));
        if (bound == INFINITY) break;                       // the whole grid has been visited
        if (GENERAL && r2 >= 0.0f) {                        // radius mode: stop once every unvisited point is farther than the radius
            if (__fmul_rn(bound, bound) > r2) break;
            continue;
        }
        if (t.kth_d < __fmul_rn(bound, bound)) break;       // strict: an equal distance with a lower index could still enter
    }
    if (lane < Kp) {
        // K > N: unfilled slots keep 0, the observable behaviour of the reference's cpp_knn_omp (knn_.cxx:59,65-67)
        out[((size_t)b * Q + out_row) * K + k0 + lane] = (t.i == 0x7fffffff) ? ((GENERAL && r2 >= 0.0f) ? -1 : 0) : (int64_t)t.i;   // radius mode pads with -1
    }
    if (!GENERAL) break;
    ld = __shfl_sync(0xffffffffu, t.d, Kp - 1);
    li = __shfl_sync(0xffffffffu, t.i, Kp - 1);
    if (li == 0x7fffffff) {                                  // the cloud is exhausted: remaining slots are 0
        for (int k = k0 + 32 + lane; k < K; k += 32) out[((size_t)b * Q + out_row) * K + k] = r2 >= 0.0f ? -1 : 0;
        break;
    }
    }
}

// ---------------------------------------------------------------------------------------------- coverage sampler
// cpp_knn_batch_distance_pick (knn_.cxx:138-203): per cloud, nqueries times: among the points whose coverage counter `used` equals
// the current level pick one uniformly (mt19937 draw % count; the reference seeds with time(0), here the seed is an argument),
// take its K nearest neighbours, add 1 to the counter of every neighbour and 100 to the picked point's; when no point is left at
// the current level the level becomes min(used).  Inherently sequential per cloud: one CTA per cloud walks the queries; the scan for
// the r-th candidate, the brute-force distance pass and the top-K merge are CTA-parallel.  The single RNG stream of the reference's
// serial loop is reproduced exactly: cloud b discards the b·nqueries draws that the clouds before it consume (one draw per query).
struct Mt19937 {
    unsigned s[624];
    int idx;
    __device__ void seed(unsigned v) {
        s[0] = v;
        for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (unsigned)i;
        idx = 624;
    }
    __device__ unsigned next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const unsigned y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
                s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        unsigned y = s[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

constexpr int kPickThreads = 1024;
__global__ void __launch_bounds__(kPickThreads) distance_pick_kernel(const float* __restrict__ pts, int* __restrict__ used_all,
                                                                     int64_t* __restrict__ out_idx, float* __restrict__ out_q, int N, int Q,
                                                                     int K, unsigned seed) {
    __shared__ Mt19937 rng;
    __shared__ int s_cnt[kPickThreads / 32], s_min[kPickThreads / 32];
    __shared__ int s_total, s_level, s_pick, s_rank;
    __shared__ float s_cd[32 * 32];
    __shared__ int s_ci[32 * 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* P = pts + (size_t)b * N * 3;
    int* used = used_all + (size_t)b * N;
    const int chunk = (N + kPickThreads - 1) / kPickThreads;            // contiguous chunk per thread: candidates keep index order
    const int c0 = min(tid * chunk, N), c1 = min(c0 + chunk, N);
    for (int i = tid; i < N; i += kPickThreads) used[i] = 0;
    if (tid == 0) {
        rng.seed(seed);
        for (long long i = 0; i < (long long)b * Q; ++i) rng.next();
        s_level = 0;
    }
    __syncthreads();
    for (int q = 0; q < Q; ++q) {
        // ---- candidates: points with used == level (ascending index); if none, level = min(used)
        int cnt, pre;
        while (true) {
            const int level = s_level;
            cnt = 0;
            int mn = 0x7fffffff;
            for (int i = c0; i < c1; ++i) {
                const int u = used[i];
                cnt += (u == level);
                mn = min(mn, u);
            }
            int inc = cnt;                                             // inclusive warp scan of the counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (lane == 31) s_cnt[warp] = inc;
            if (lane == 0) s_min[warp] = mn;
            __syncthreads();
            if (warp == 0) {
                int w = s_cnt[lane], m = s_min[lane];
                int winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc += v;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                s_cnt[lane] = winc - w;                                // exclusive warp offsets
                if (lane == 31) s_total = winc;
                if (lane == 0 && winc == 0) {}                         // (total read below)
                if (lane == 0) s_min[0] = m;
            }
            __syncthreads();
            pre = s_cnt[warp] + inc - cnt;                             // exclusive prefix of this thread
            if (s_total > 0) break;
            __syncthreads();
            if (tid == 0) s_level = s_min[0];
            __syncthreads();
        }
        if (tid == 0) s_rank = (int)((unsigned long long)rng.next() % (unsigned long long)s_total);
        __syncthreads();
        const int rank = s_rank;
        if (rank >= pre && rank < pre + cnt) {                         // exactly one thread owns the rank-th candidate
            const int level = s_level;
            int seen = pre;
            for (int i = c0; i < c1; ++i)
                if (used[i] == level && seen++ == rank) { s_pick = i; break; }
        }
        __syncthreads();
        const int pick = s_pick;
        const float qx = __ldg(P + 3 * (size_t)pick), qy = __ldg(P + 3 * (size_t)pick + 1), qz = __ldg(P + 3 * (size_t)pick + 2);
        // ---- K nearest neighbours of the picked point: every warp keeps the top-K of its stripe, warp 0 merges the 32 lists
        TopK t;
        t.d = INFINITY; t.i = 0x7fffffff; t.kth_d = INFINITY; t.kth_i = 0x7fffffff;
        for (int base = warp * 32; base < N; base += kPickThreads) {
            const int p = base + lane;
            float d = INFINITY;
            int idx = 0x7fffffff;
            if (p < N) {
                d = sqdist(qx, qy, qz, __ldg(P + 3 * (size_t)p), __ldg(P + 3 * (size_t)p + 1), __ldg(P + 3 * (size_t)p + 2));
                idx = p;
            }
            const bool pass = (d < t.kth_d) || (d == t.kth_d && idx < t.kth_i);
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float dc = __shfl_sync(0xffffffffu, d, src);
                const int ic = __shfl_sync(0xffffffffu, idx, src);
                if (!((dc < t.kth_d) || (dc == t.kth_d && ic < t.kth_i))) continue;
                const bool less = (t.d < dc) || (t.d == dc && t.i < ic);
                const int pos = __popc(__ballot_sync(0xffffffffu, less));
                const float ud = __shfl_up_sync(0xffffffffu, t.d, 1);
                const int ui = __shfl_up_sync(0xffffffffu, t.i, 1);
                if (lane == pos) { t.d = dc; t.i = ic; }
                else if (lane > pos) { t.d = ud; t.i = ui; }
                t.kth_d = __shfl_sync(0xffffffffu, t.d, K - 1);
                t.kth_i = __shfl_sync(0xffffffffu, t.i, K - 1);
            }
        }
        s_cd[warp * 32 + lane] = t.d;
        s_ci[warp * 32 + lane] = t.i;
        __syncthreads();
        if (warp == 0) {
            TopK r;
            r.d = INFINITY; r.i = 0x7fffffff; r.kth_d = INFINITY; r.kth_i = 0x7fffffff;
            for (int w = 0; w < kPickThreads / 32; ++w) {
                const float d = lane < K ? s_cd[w * 32 + lane] : INFINITY;
                const int idx = lane < K ? s_ci[w * 32 + lane] : 0x7fffffff;
                const bool pass = (d < r.kth_d) || (d == r.kth_d && idx < r.kth_i);
                unsigned m = __ballot_sync(0xffffffffu, pass);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float dc = __shfl_sync(0xffffffffu, d, src);
                    const int ic = __shfl_sync(0xffffffffu, idx, src);
                    if (!((dc < r.kth_d) || (dc == r.kth_d && ic < r.kth_i))) continue;
                    const bool less = (r.d < dc) || (r.d == dc && r.i < ic);
                    const int pos = __popc(__ballot_sync(0xffffffffu, less));
                    const float ud = __shfl_up_sync(0xffffffffu, r.d, 1);
                    const int ui = __shfl_up_sync(0xffffffffu, r.i, 1);
                    if (lane == pos) { r.d = dc; r.i = ic; }
                    else if (lane > pos) { r.d = ud; r.i = ui; }
                    r.kth_d = __shfl_sync(0xffffffffu, r.d, K - 1);
                    r.kth_i = __shfl_sync(0xffffffffu, r.i, K - 1);
                }
            }
            if (lane < K) {
                const bool have = r.i != 0x7fffffff;
                out_idx[((size_t)b * Q + q) * K + lane] = have ? (int64_t)r.i : 0;
                if (have) used[r.i] += 1;                              // neighbour ids are distinct: no conflicts
            }
            __syncwarp();
            if (lane == 0) {
                used[pick] += 100;
                float* oq = out_q + ((size_t)b * Q + q) * 3;
                oq[0] = qx; oq[1] = qy; oq[2] = qz;
            }
        }
        __syncthreads();
    }
}

}  // namespace knn
}  // namespace crf

using namespace crf;

extern "C" {

size_t crfconv_knn_workspace_bytes(int64_t B, int64_t N, int64_t Q, int64_t K) {
    (void)Q; (void)K;
    const size_t cap = (size_t)knn::cell_cap(N);
    size_t tot = 0;
    tot += align_up(B * sizeof(knn::Grid), 256);
    tot += align_up(B * 6 * sizeof(unsigned), 256);
    tot += align_up((size_t)B * N * sizeof(int), 256);            // cell_of
    tot += align_up((size_t)B * (cap + 1) * sizeof(int), 256);    // cell_count / cell_start
    tot += align_up((size_t)B * cap * sizeof(int), 256);          // cursor
    tot += align_up((size_t)B * N * sizeof(float4), 256);         // sorted
    tot += align_up((size_t)B * 2 * knn::kEdgeStride * sizeof(float), 256);
    return tot;
}

// Device-pointer form of cpp_knn_batch / cpp_knn_batch_omp (knn_.h:13-19).  dim is fixed to 3.
static int knn_search(const float* pts, int64_t B, int64_t N, const float* queries, int64_t Q, int64_t K, int64_t* out_idx, void* workspace,
                      size_t workspace_bytes, void* stream_, float r2);

int crfconv_knn_batch(const float* pts, int64_t B, int64_t N, const float* queries, int64_t Q, int64_t K,
                      int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream_) {
    return knn_search(pts, B, N, queries, Q, K, out_idx, workspace, workspace_bytes, stream_, -1.0f);
}

// Radius search on the same grid: out_idx [B,Q,K] receives, per query, the (up to) K smallest indices of the support points with
// squared distance <= r·r, ascending, padded with -1 — torch_cluster.radius(x, y, r, max_num_neighbors = K) on equal-sized clouds
// (what models/continuous_crf_conv.py:52, discrete_crf_conv.py:44 and point_conv.py:150,180 build their graphs with).
int crfconv_radius_batch(const float* pts, int64_t B, int64_t N, const float* queries, int64_t Q, float r, int64_t K, int64_t* out_idx,
                         void* workspace, size_t workspace_bytes, void* stream_) {
    if (!(r >= 0.0f)) return CRF_ERR_INVALID_ARG;
    return knn_search(pts, B, N, queries, Q, K, out_idx, workspace, workspace_bytes, stream_, r * r);
}

static int knn_search(const float* pts, int64_t B, int64_t N, const float* queries, int64_t Q, int64_t K, int64_t* out_idx, void* workspace,
                      size_t workspace_bytes, void* stream_, float r2) {
    if (B < 0 || N < 0 || Q < 0 || K < 0) return CRF_ERR_INVALID_ARG;
    if (B == 0 || Q == 0 || K == 0) return CRF_OK;
    if (N == 0) return CRF_ERR_INVALID_ARG;                          // the reference asserts npts != 0 (KDTreeTableAdaptor.h:136)
    if (K > 1024) return CRF_ERR_UNSUPPORTED;
    if (N > (1 << 30) || Q > (1 << 26) * 8 || B > 65535) return CRF_ERR_INVALID_ARG;
    if (!pts || !queries || !out_idx || !workspace) return CRF_ERR_INVALID_ARG;
    if (workspace_bytes < crfconv_knn_workspace_bytes(B, N, Q, K)) return CRF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream_;
    const int cap = (int)knn::cell_cap(N);

    Carver cv(workspace);
    knn::Grid* grids = cv.take<knn::Grid>(B);
    unsigned* bbox = cv.take<unsigned>(B * 6);
    int* cell_of = cv.take<int>((size_t)B * N);
    int* cell_start = cv.take<int>((size_t)B * (cap + 1));
    int* cursor = cv.take<int>((size_t)B * cap);
    float4* sorted = cv.take<float4>((size_t)B * N);
    float* edges = cv.take<float>((size_t)B * 2 * knn::kEdgeStride);

    // target occupancy per cell ~ K/4: the K-th neighbour then sits at ≈ 1 cell, so shell r=1 usually suffices
    float occ = fminf(fmaxf((float)K * 0.25f, 1.0f), 8.0f);

    knn::init_kernel<<<(int)ceil_div(B * 6, 128), 128, 0, st>>>(bbox, (int)B);
    CRF_CUDA(cudaMemsetAsync(cell_start, 0, (size_t)B * (cap + 1) * sizeof(int), st));
    CRF_CUDA(cudaMemsetAsync(cursor, 0, (size_t)B * cap * sizeof(int), st));
    const int nblk = (int)std::min<int64_t>(ceil_div(N, 256), 4 * kNumSMs);
    knn::bbox_kernel<<<dim3(nblk, (unsigned)B), 256, 0, st>>>(pts, bbox, (int)N);
    knn::setup_kernel<<<(unsigned)B, 256, 0, st>>>(bbox, grids, edges, (int)N, occ);
    knn::count_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)B), 256, 0, st>>>(pts, grids, cell_of, cell_start, (int)N, cap);
    knn::scan_kernel<<<(unsigned)B, 1024, 0, st>>>(cell_start, grids, cap);
    knn::scatter_kernel<<<dim3((unsigned)ceil_div(N, 256), (unsigned)B), 256, 0, st>>>(pts, cell_of, cell_start, cursor, sorted, (int)N, cap);
    const bool self = queries == pts && Q == N, general = K > 32 || r2 >= 0.0f;
    const dim3 qgrid((unsigned)ceil_div(Q, 8), (unsigned)B);
#define CRF_KNN_LAUNCH(S, G) knn::query_kernel<S, G><<<qgrid, 256, 0, st>>>(sorted, cell_start, grids, edges, queries, out_idx, (int)N, (int)Q, (int)K, cap, r2)
    if (self && !general) CRF_KNN_LAUNCH(true, false);
    else if (self) CRF_KNN_LAUNCH(true, true);
    else if (!general) CRF_KNN_LAUNCH(false, false);
    else CRF_KNN_LAUNCH(false, true);
#undef CRF_KNN_LAUNCH
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

// Device-pointer form of cpp_knn_batch_distance_pick / _omp (knn_.cxx:138-271, knn.pyx:111-149) with the RNG seed as an argument
// (the reference seeds std::mt19937 with time(0)).  pts [B,N,3]; out_idx [B,Q,K] i64; out_queries [B,Q,3] f32 (the picked points);
// workspace: B·N int32 coverage counters.  K <= 32.  One CTA per cloud (the algorithm is sequential in the queries).
size_t crfconv_knn_distance_pick_workspace_bytes(int64_t B, int64_t N) { return align_up((size_t)B * N * sizeof(int), 256); }

int crfconv_knn_batch_distance_pick(const float* pts, int64_t B, int64_t N, int64_t nqueries, int64_t K, uint32_t seed, int64_t* out_idx,
                                    float* out_queries, void* workspace, size_t workspace_bytes, void* stream_) {
    if (B < 0 || N <= 0 || nqueries < 0 || K <= 0) return CRF_ERR_INVALID_ARG;
    if (B == 0 || nqueries == 0) return CRF_OK;
    if (K > 32) return CRF_ERR_UNSUPPORTED;
    if (!pts || !out_idx || !out_queries || !workspace || N > (1 << 30) || B > 65535) return CRF_ERR_INVALID_ARG;
    if (workspace_bytes < crfconv_knn_distance_pick_workspace_bytes(B, N)) return CRF_ERR_WORKSPACE;
    knn::distance_pick_kernel<<<(unsigned)B, knn::kPickThreads, 0, (cudaStream_t)stream_>>>(pts, (int*)workspace, out_idx, out_queries, (int)N,
                                                                                        (int)nqueries, (int)K, seed);
    CRF_LAUNCH_CHECK();
    return CRF_OK;
}

// HOST pointers — the drop-in for `void cpp_knn_batch_distance_pick(const float* batch_data, size_t batch_size, size_t npts, size_t dim,
// float* batch_queries, size_t nqueries, size_t K, long* batch_indices)` (knn_.h) plus the seed.
int crfconv_cpp_knn_batch_distance_pick(const float* batch_data, size_t batch_size, size_t npts, size_t dim, float* batch_queries,
                                        size_t nqueries, size_t K, int64_t* batch_indices, uint32_t seed) {
    if (dim != 3) return CRF_ERR_UNSUPPORTED;
    if (batch_size == 0 || nqueries == 0) return CRF_OK;
    if (!batch_data || !batch_queries || !batch_indices || npts == 0) return CRF_ERR_INVALID_ARG;
    float *d_pts = nullptr, *d_q = nullptr;
    int64_t* d_idx = nullptr;
    void* ws = nullptr;
    const size_t wsb = crfconv_knn_distance_pick_workspace_bytes((int64_t)batch_size, (int64_t)npts);
    int rc = CRF_OK;
    auto fail = [&](cudaError_t e) { rc = (int)e; };
    cudaError_t e;
    if ((e = cudaMalloc(&d_pts, batch_size * npts * 3 * sizeof(float))) != cudaSuccess) fail(e);
    if (rc == CRF_OK && (e = cudaMalloc(&d_q, batch_size * nqueries * 3 * sizeof(float))) != cudaSuccess) fail(e);
    if (rc == CRF_OK && (e = cudaMalloc(&d_idx, batch_size * nqueries * K * sizeof(int64_t))) != cudaSuccess) fail(e);
    if (rc == CRF_OK && (e = cudaMalloc(&ws, wsb)) != cudaSuccess) fail(e);
    if (rc == CRF_OK && (e = cudaMemcpy(d_pts, batch_data, batch_size * npts * 3 * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) fail(e);
    if (rc == CRF_OK)
        rc = crfconv_knn_batch_distance_pick(d_pts, (int64_t)batch_size, (int64_t)npts, (int64_t)nqueries, (int64_t)K, seed, d_idx, d_q, ws, wsb, nullptr);
    if (rc == CRF_OK && (e = cudaMemcpy(batch_indices, d_idx, batch_size * nqueries * K * sizeof(int64_t), cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e);
    if (rc == CRF_OK && (e = cudaMemcpy(batch_queries, d_q, batch_size * nqueries * 3 * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) fail(e);
    cudaFree(d_pts); cudaFree(d_q); cudaFree(d_idx); cudaFree(ws);
    return rc;
}

}  // extern "C"
