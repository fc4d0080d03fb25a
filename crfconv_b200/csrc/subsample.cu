// Voxel-grid (barycentre) subsampling on sm_100a — B200-native replacement of the reference's
// grid_subsampling core (/root/reference/utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106,
// grid_subsampling.h:10-80, cpp_utils/cloud/cloud.cpp:27-66; arithmetic spec in SURVEY.md Appendix A.2).
//
// The reference walks the points once, accumulating into an unordered_map keyed by the voxel index.  Float sums are
// order dependent, so bit-exactness requires every voxel to be summed in ORIGINAL POINT ORDER.  Device pipeline:
//   min/max → origin / NX / NY (exact f32 ops, explicit _rn intrinsics, true division) → 64-bit voxel key per point →
//   stable LSD radix sort of (key, original index) (hand-written, 8-bit digits, warp match_any ranking) →
//   segment heads + scan → one thread per (voxel, channel) sums its segment sequentially (original order preserved by
//   the stable sort) → barycentre / mean feature; label vote with ≤8 register slots per voxel.
// Two things are genuinely host-defined in the reference and are reproduced on the host with the same container:
//   * row order = libstdc++ std::unordered_map<size_t,…> iteration order (.cpp:48,85) — optional (order=1): the
//     first-seen key sequence is replayed through a std::unordered_map; order=0 emits ascending key order.
//   * label-vote ties = first maximum in std::unordered_map<int,int> iteration order (.cpp:100-101) — only tied /
//     overflowing voxels are replayed.
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "../../include/crfconv_b200.h"
#include "common.cuh"
#include "scan.cuh"

namespace crf {
namespace gs {

struct Params {
    float org[3];
    float dl;
    unsigned long long NX, NY, NZ;
    float mn[3], mx[3];
};

__global__ void init_kernel(unsigned* bbox) {
    if (threadIdx.x < 6) bbox[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ p, unsigned* bbox, int64_t N) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __ldg(p + 3 * i + c);
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
            atomicMin(bbox + c, f2ord(mn[c]));
            atomicMax(bbox + 3 + c, f2ord(mx[c]));
        }
    }
}

// grid_subsampling.cpp:27-31.  float→size_t of a negative value is undefined in the reference; here it saturates to 0.
__device__ __forceinline__ unsigned long long vox_coord(float x, float org, float dl) {
    return (unsigned long long)floorf(__fdiv_rn(__fsub_rn(x, org), dl));
}

__global__ void params_kernel(const unsigned* __restrict__ bbox, Params* P, float dl) {
    if (threadIdx.x != 0) return;
    Params p;
    p.dl = dl;
    const float inv = __fdiv_rn(1.0f, dl);
    for (int c = 0; c < 3; ++c) {
        p.mn[c] = ord2f(bbox[c]);
        p.mx[c] = ord2f(bbox[3 + c]);
        p.org[c] = __fmul_rn(floorf(__fmul_rn(p.mn[c], inv)), dl);
    }
    p.NX = vox_coord(p.mx[0], p.org[0], dl) + 1;
    p.NY = vox_coord(p.mx[1], p.org[1], dl) + 1;
    p.NZ = vox_coord(p.mx[2], p.org[2], dl) + 1;   // not used by the key (as in the reference); bounds the key width
    *P = p;
}

__global__ void __launch_bounds__(256) key_kernel(const float* __restrict__ pts, const Params* __restrict__ P,
                                                  unsigned long long* __restrict__ keys, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const Params& p = *P;
    const unsigned long long ix = vox_coord(__ldg(pts + 3 * i), p.org[0], p.dl);
    const unsigned long long iy = vox_coord(__ldg(pts + 3 * i + 1), p.org[1], p.dl);
    const unsigned long long iz = vox_coord(__ldg(pts + 3 * i + 2), p.org[2], p.dl);
    keys[i] = ix + p.NX * iy + p.NX * p.NY * iz;   // .cpp:56
}

// ------------------------------------------------------------------------------- stable LSD radix sort
constexpr int kRsThreads = 256, kRsWarps = 8, kRsIters = 16, kRsTile = kRsThreads * kRsIters;

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const unsigned long long* __restrict__ keys,
                                                             int* __restrict__ counts, int64_t n, int shift, int nblocks) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll 4
    for (int it = 0; it < kRsIters; ++it) {
        const int64_t i = base + it * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 255ull)], 1);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major ⇒ one flat scan gives offsets
}

__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in,
                                                                unsigned long long* __restrict__ keys_out,
                                                                int* __restrict__ vals_out,
                                                                const int* __restrict__ offsets, int64_t n, int shift,
                                                                int nblocks) {
    __shared__ int cnt[kRsWarps][256];
    for (int j = threadIdx.x; j < kRsWarps * 256; j += kRsThreads) (&cnt[0][0])[j] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)w * (kRsIters * 32);
    unsigned long long k[kRsIters];
    int v[kRsIters], rk[kRsIters];
#pragma unroll
    for (int it = 0; it < kRsIters; ++it) {
        const int64_t i = wbase + it * 32 + lane;
        const bool valid = i < n;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        rk[it] = 0; k[it] = 0; v[it] = 0;
        if (valid) {
            k[it] = keys_in[i];
            v[it] = vals_in ? vals_in[i] : (int)i;
            const int d = (int)((k[it] >> shift) & 255ull);
            const unsigned m = __match_any_sync(act, d);
            const int prior = cnt[w][d];
            __syncwarp(act);
            if (lane == __ffs(m) - 1) cnt[w][d] = prior + __popc(m);
            __syncwarp(act);
            rk[it] = prior + __popc(m & lt);          // stable: lower lanes (= lower positions) first
        }
    }
    __syncthreads();
    {   // thread d: exclusive prefix over the block's warps, seeded with the global offset of (digit d, this block)
        const int d = threadIdx.x;
        int run = offsets[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < kRsWarps; ++ww) {
            const int t = cnt[ww][d];
            cnt[ww][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kRsIters; ++it) {
        const int64_t i = wbase + it * 32 + lane;
        if (i < n) {
            const int d = (int)((k[it] >> shift) & 255ull);
            const int pos = cnt[w][d] + rk[it];
            keys_out[pos] = k[it];
            vals_out[pos] = v[it];
        }
    }
}

// ------------------------------------------------------------------------------------------ segments
__global__ void __launch_bounds__(256) head_kernel(const unsigned long long* __restrict__ keys, int* __restrict__ head, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) segstart_kernel(const int* __restrict__ head, const int* __restrict__ vex,
                                                       int* __restrict__ seg_start, int* __restrict__ M_out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (head[i]) seg_start[vex[i]] = (int)i;
    if (i == n - 1) {
        const int M = vex[i] + head[i];
        seg_start[M] = (int)n;
        *M_out = M;
    }
}

// first-seen order: voxel v was first touched by original point f_v = idx_sorted[seg_start[v]] (stable sort)
__global__ void __launch_bounds__(256) mark_first_kernel(const int* __restrict__ seg_start, const int* __restrict__ idx_sorted,
                                                         int* __restrict__ mark, int M) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < M) mark[idx_sorted[seg_start[v]]] = 1;
}

__global__ void __launch_bounds__(256) firstseen_kernel(const int* __restrict__ seg_start, const int* __restrict__ idx_sorted,
                                                        const unsigned long long* __restrict__ keys_sorted,
                                                        const int* __restrict__ rank, unsigned long long* __restrict__ fs_key,
                                                        int* __restrict__ fs_vox, int M) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= M) return;
    const int s = seg_start[v];
    const int j = rank[idx_sorted[s]];
    fs_key[j] = keys_sorted[s];
    fs_vox[j] = v;
}

__global__ void __launch_bounds__(256) rowmap_kernel(const int* __restrict__ fs_vox, const int* __restrict__ row_of_rank,
                                                     int* __restrict__ row_of_vox, int M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) row_of_vox[fs_vox[j]] = row_of_rank[j];
}

__global__ void __launch_bounds__(256) voxkey_kernel(const int* __restrict__ seg_start, const unsigned long long* __restrict__ keys_sorted,
                                                     const int* __restrict__ row_of_vox, unsigned long long* __restrict__ out_keys, int M) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < M) out_keys[row_of_vox ? row_of_vox[v] : v] = keys_sorted[seg_start[v]];
}

// ------------------------------------------------------------------------------------------ reductions
// One thread per (voxel, channel); channels 0..2 = xyz, 3.. = features.  Sequential f32 adds in original order.
__global__ void __launch_bounds__(256) reduce_kernel(const float* __restrict__ pts, const float* __restrict__ feats, int fdim,
                                                     const int* __restrict__ seg_start, const int* __restrict__ idx_sorted,
                                                     const int* __restrict__ row_of_vox, float* __restrict__ out_pts,
                                                     float* __restrict__ out_feats, int M) {
    const int nch = 3 + fdim;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)M * nch) return;
    const int v = (int)(t / nch), ch = (int)(t % nch);
    const int s0 = seg_start[v], s1 = seg_start[v + 1];
    float acc = 0.0f;
    if (ch < 3) {
        for (int s = s0; s < s1; ++s) acc = __fadd_rn(acc, __ldg(pts + 3 * (int64_t)idx_sorted[s] + ch));
    } else {
        const int f = ch - 3;
        for (int s = s0; s < s1; ++s) acc = __fadd_rn(acc, __ldg(feats + (int64_t)idx_sorted[s] * fdim + f));
    }
    const int count = s1 - s0;
    const int64_t row = row_of_vox ? row_of_vox[v] : v;
    if (ch < 3) out_pts[row * 3 + ch] = __fmul_rn(acc, (float)(1.0 / (double)count));     // .cpp:87 PointXYZ * (float)(1.0/count)
    else        out_feats[row * fdim + ch - 3] = __fdiv_rn(acc, (float)count);            // .cpp:91-94
}

// One thread per (voxel, label dim): plurality vote with 8 register slots.  tie[...] = 1 when the winner is not unique
// or more than 8 distinct labels occur; those voxels are replayed on the host with std::unordered_map<int,int>.
__global__ void __launch_bounds__(256) vote_kernel(const int* __restrict__ cls, int ldim, const int* __restrict__ seg_start,
                                                   const int* __restrict__ idx_sorted, const int* __restrict__ row_of_vox,
                                                   int* __restrict__ out_cls, unsigned char* __restrict__ tie, int M) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)M * ldim) return;
    const int v = (int)(t / ldim), l = (int)(t % ldim);
    const int s0 = seg_start[v], s1 = seg_start[v + 1];
    int lab[8], cnt[8], ns = 0;
    bool overflow = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) { lab[j] = 0; cnt[j] = 0; }
    for (int s = s0; s < s1; ++s) {
        const int L = __ldg(cls + (int64_t)idx_sorted[s] * ldim + l);
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < ns && lab[j] == L) { cnt[j] += 1; found = true; }
        if (!found) {
            if (ns < 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j == ns) { lab[j] = L; cnt[j] = 1; }
                ns += 1;
            } else {
                overflow = true;
            }
        }
    }
    int best = -1, best_lab = 0, nbest = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < ns) {
            if (cnt[j] > best) { best = cnt[j]; best_lab = lab[j]; nbest = 1; }
            else if (cnt[j] == best) nbest += 1;
        }
    }
    const int64_t row = row_of_vox ? row_of_vox[v] : v;
    out_cls[row * ldim + l] = best_lab;
    tie[(int64_t)v * ldim + l] = (overflow || nbest > 1) ? 1 : 0;
}

__global__ void __launch_bounds__(256) patch_kernel(int* __restrict__ out_cls, const long long* __restrict__ where,
                                                    const int* __restrict__ what, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out_cls[where[i]] = what[i];
}

inline int key_bits(unsigned long long maxkey) {
    int b = 0;
    while (maxkey) { ++b; maxkey >>= 1; }
    return std::max(b, 1);
}

}  // namespace gs
}  // namespace crf

using namespace crf;

extern "C" {

size_t crfconv_grid_subsample_workspace_bytes(int64_t N, int64_t fdim, int64_t ldim) {
    (void)fdim;
    const size_t n = (size_t)std::max<int64_t>(N, 1);
    const size_t nblocks = (size_t)ceil_div((int64_t)n, gs::kRsTile);
    size_t tot = 0;
    tot += 2 * align_up(n * 8, 256);                       // keys ping/pong
    tot += 2 * align_up(n * 4, 256);                       // vals ping/pong
    tot += align_up(256 * nblocks * 4, 256);               // digit counts
    tot += align_up(scan::workspace_ints((int64_t)std::max(n, 256 * nblocks)) * 4, 256);
    tot += 2 * align_up(n * 4, 256);                       // head / vex (re-used: mark / rank)
    tot += align_up((n + 1) * 4, 256);                     // seg_start
    tot += 2 * align_up(n * 4, 256);                       // row_of_vox, fs_vox / row_of_rank
    tot += align_up(n * 4, 256);                           // row_of_rank
    tot += align_up(n * 8, 256);                           // fs_key
    tot += align_up(n * (size_t)std::max<int64_t>(ldim, 1), 256);   // tie flags
    tot += align_up(n * (size_t)std::max<int64_t>(ldim, 1) * 12, 256);   // patches (where: 8 B, what: 4 B)
    tot += 1024;                                           // bbox, params, M
    return tot;
}

// Device-pointer grid subsampling.  points [N,3] f32, features [N,fdim] f32 or NULL, classes [N,ldim] i32 or NULL.
// Outputs must hold N rows.  order: 0 = ascending voxel key, 1 = reference (libstdc++ unordered_map) row order.
// out_keys (optional, device, N × u64) receives the voxel key of every output row.  *M_out (host) = number of rows.
// Synchronises the stream (the row count is data dependent, like the reference's return value).
int crfconv_grid_subsample(const float* points, int64_t N, const float* features, int64_t fdim, const int32_t* classes,
                           int64_t ldim, float sampleDl, int order, float* out_points, float* out_features,
                           int32_t* out_classes, unsigned long long* out_keys, int64_t* M_out, void* workspace,
                           size_t workspace_bytes, void* stream_) {
    if (!M_out) return CRF_ERR_INVALID_ARG;
    *M_out = 0;
    if (N < 0 || fdim < 0 || ldim < 0 || N > 0x7fffffff || !(sampleDl > 0.0f)) return CRF_ERR_INVALID_ARG;
    if (N == 0) return CRF_OK;
    if (!points || !out_points || !workspace) return CRF_ERR_INVALID_ARG;
    if ((fdim > 0 && (!features || !out_features)) || (ldim > 0 && (!classes || !out_classes))) return CRF_ERR_INVALID_ARG;
    if (workspace_bytes < crfconv_grid_subsample_workspace_bytes(N, fdim, ldim)) return CRF_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream_;
    const int nblocks = (int)ceil_div(N, gs::kRsTile);
    const int64_t ld = std::max<int64_t>(ldim, 1);

    Carver cv(workspace);
    unsigned long long* keysA = cv.take<unsigned long long>(N);
    unsigned long long* keysB = cv.take<unsigned long long>(N);
    int* valsA = cv.take<int>(N);
    int* valsB = cv.take<int>(N);
    int* counts = cv.take<int>((size_t)256 * nblocks);
    int* scan_ws = cv.take<int>(scan::workspace_ints(std::max<int64_t>(N, (int64_t)256 * nblocks)));
    int* head = cv.take<int>(N);
    int* vex = cv.take<int>(N);
    int* seg_start = cv.take<int>(N + 1);
    int* row_of_vox = cv.take<int>(N);
    int* fs_vox = cv.take<int>(N);
    int* row_of_rank = cv.take<int>(N);
    unsigned long long* fs_key = cv.take<unsigned long long>(N);
    unsigned char* tie = cv.take<unsigned char>((size_t)N * ld);
    long long* patch_where = cv.take<long long>((size_t)N * ld);
    int* patch_what = cv.take<int>((size_t)N * ld);
    unsigned* bbox = cv.take<unsigned>(8);
    gs::Params* dparams = cv.take<gs::Params>(1);
    int* dM = cv.take<int>(1);

    // ---- 1. bounding box → grid origin and dimensions
    gs::init_kernel<<<1, 32, 0, st>>>(bbox);
    gs::minmax_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N, 256), 8 * kNumSMs), 256, 0, st>>>(points, bbox, N);
    gs::params_kernel<<<1, 32, 0, st>>>(bbox, dparams, sampleDl);
    gs::Params hp;
    CRF_CUDA(cudaMemcpyAsync(&hp, dparams, sizeof(hp), cudaMemcpyDeviceToHost, st));
    CRF_CUDA(cudaStreamSynchronize(st));
    // widest key = NX*NY*NZ - 1 (saturating)
    unsigned long long maxkey = ~0ull;
    {
        const long double prod = (long double)hp.NX * (long double)hp.NY * (long double)hp.NZ;
        if (prod < 1.8e19L) maxkey = hp.NX * hp.NY * hp.NZ - 1ull;
    }
    const int npass = (gs::key_bits(maxkey) + 7) / 8;

    // ---- 2. keys, stable radix sort by key carrying the original index
    gs::key_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(points, dparams, keysA, N);
    unsigned long long *kin = keysA, *kout = keysB;
    int *vin = nullptr, *vout = valsA, *vspare = valsB;
    for (int p = 0; p < npass; ++p) {
        gs::rs_hist_kernel<<<nblocks, gs::kRsThreads, 0, st>>>(kin, counts, N, 8 * p, nblocks);
        int rc = scan::exclusive(counts, counts, (int64_t)256 * nblocks, scan_ws, st);
        if (rc != CRF_OK) return rc;
        gs::rs_scatter_kernel<<<nblocks, gs::kRsThreads, 0, st>>>(kin, vin, kout, vout, counts, N, 8 * p, nblocks);
        std::swap(kin, kout);
        int* nv = vin ? vin : vspare;
        vin = vout;
        vout = nv;
    }
    const unsigned long long* keys_sorted = kin;
    const int* idx_sorted = vin;

    // ---- 3. segments (voxels) and their count
    const unsigned nb256 = (unsigned)ceil_div(N, 256);
    gs::head_kernel<<<nb256, 256, 0, st>>>(keys_sorted, head, N);
    {
        int rc = scan::exclusive(head, vex, N, scan_ws, st);
        if (rc != CRF_OK) return rc;
    }
    gs::segstart_kernel<<<nb256, 256, 0, st>>>(head, vex, seg_start, dM, N);
    int M = 0;
    CRF_CUDA(cudaMemcpyAsync(&M, dM, sizeof(int), cudaMemcpyDeviceToHost, st));
    CRF_CUDA(cudaStreamSynchronize(st));
    const unsigned mb256 = (unsigned)ceil_div(M, 256);

    // ---- 4. optional reference row order (host replay of the first-seen key sequence)
    const int* rows = nullptr;
    if (order == 1) {
        int* mark = head;   // head / vex are free again
        int* rank = vex;
        CRF_CUDA(cudaMemsetAsync(mark, 0, (size_t)N * sizeof(int), st));
        gs::mark_first_kernel<<<mb256, 256, 0, st>>>(seg_start, idx_sorted, mark, M);
        int rc = scan::exclusive(mark, rank, N, scan_ws, st);
        if (rc != CRF_OK) return rc;
        gs::firstseen_kernel<<<mb256, 256, 0, st>>>(seg_start, idx_sorted, keys_sorted, rank, fs_key, fs_vox, M);
        std::vector<unsigned long long> hkeys((size_t)M);
        CRF_CUDA(cudaMemcpyAsync(hkeys.data(), fs_key, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
        CRF_CUDA(cudaStreamSynchronize(st));
        std::unordered_map<size_t, int> replay;             // the reference's container type (.cpp:48)
        for (int j = 0; j < M; ++j) replay.emplace((size_t)hkeys[j], j);
        std::vector<int> hrow((size_t)M);
        int r = 0;
        for (auto& kv : replay) hrow[kv.second] = r++;
        CRF_CUDA(cudaMemcpyAsync(row_of_rank, hrow.data(), (size_t)M * 4, cudaMemcpyHostToDevice, st));
        gs::rowmap_kernel<<<mb256, 256, 0, st>>>(fs_vox, row_of_rank, row_of_vox, M);
        CRF_CUDA(cudaStreamSynchronize(st));                 // hrow must outlive the copy
        rows = row_of_vox;
    }

    // ---- 5. per-voxel sequential sums, barycentres, mean features, label votes
    gs::reduce_kernel<<<(unsigned)ceil_div((int64_t)M * (3 + fdim), 256), 256, 0, st>>>(
        points, features, (int)fdim, seg_start, idx_sorted, rows, out_points, out_features, M);
    if (out_keys) gs::voxkey_kernel<<<mb256, 256, 0, st>>>(seg_start, keys_sorted, rows, out_keys, M);
    if (ldim > 0) {
        gs::vote_kernel<<<(unsigned)ceil_div((int64_t)M * ldim, 256), 256, 0, st>>>(classes, (int)ldim, seg_start, idx_sorted,
                                                                                rows, out_classes, tie, M);
        std::vector<unsigned char> htie((size_t)M * ldim);
        CRF_CUDA(cudaMemcpyAsync(htie.data(), tie, htie.size(), cudaMemcpyDeviceToHost, st));
        CRF_CUDA(cudaStreamSynchronize(st));
        bool any = false;
        for (unsigned char t : htie) if (t) { any = true; break; }
        if (any) {
            std::vector<int> hseg((size_t)M + 1), hidx((size_t)N), hcls((size_t)N * ldim), hrows;
            CRF_CUDA(cudaMemcpyAsync(hseg.data(), seg_start, hseg.size() * 4, cudaMemcpyDeviceToHost, st));
            CRF_CUDA(cudaMemcpyAsync(hidx.data(), idx_sorted, hidx.size() * 4, cudaMemcpyDeviceToHost, st));
            CRF_CUDA(cudaMemcpyAsync(hcls.data(), classes, hcls.size() * 4, cudaMemcpyDeviceToHost, st));
            if (rows) {
                hrows.resize((size_t)M);
                CRF_CUDA(cudaMemcpyAsync(hrows.data(), rows, hrows.size() * 4, cudaMemcpyDeviceToHost, st));
            }
            CRF_CUDA(cudaStreamSynchronize(st));
            std::vector<long long> where;
            std::vector<int> what;
            for (int v = 0; v < M; ++v)
                for (int l = 0; l < (int)ldim; ++l) {
                    if (!htie[(size_t)v * ldim + l]) continue;
                    std::unordered_map<int, int> hist;       // the reference's container type (grid_subsampling.h:20)
                    for (int s = hseg[v]; s < hseg[v + 1]; ++s) hist[hcls[(size_t)hidx[s] * ldim + l]] += 1;
                    int best = -1, best_lab = 0;
                    for (auto& kv : hist)
                        if (kv.second > best) { best = kv.second; best_lab = kv.first; }   // first maximum in iteration order
                    where.push_back((long long)(rows ? hrows[v] : v) * ldim + l);
                    what.push_back(best_lab);
                }
            const int np = (int)where.size();
            CRF_CUDA(cudaMemcpyAsync(patch_where, where.data(), (size_t)np * 8, cudaMemcpyHostToDevice, st));
            CRF_CUDA(cudaMemcpyAsync(patch_what, what.data(), (size_t)np * 4, cudaMemcpyHostToDevice, st));
            gs::patch_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, st>>>(out_classes, patch_where, patch_what, np);
            CRF_CUDA(cudaStreamSynchronize(st));
        }
    }
    CRF_LAUNCH_CHECK();
    CRF_CUDA(cudaStreamSynchronize(st));
    *M_out = M;
    return CRF_OK;
}

}  // extern "C"
