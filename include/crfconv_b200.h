/* crfconv_b200 — C ABI of the B200 (sm_100a) hot path of CRFConv.
 *
 * Plain pointers and sizes only: no torch, numpy or C++ types cross this boundary.  Every entry point returns an
 * `int` status (0 = ok, <0 = CRF_ERR_* argument / support errors, >0 = cudaError_t) and never throws.
 * "device" pointers are CUDA device memory of the current device; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  Device entry points are stream-ordered and do not synchronise unless stated.
 * Each declaration cites the reference interface it stands in for (paths relative to the reference repository
 * yangfei1223/CRFConv).  INTEGRATION.md shows the binding a reference maintainer would add.
 */
#ifndef CRFCONV_B200_H_
#define CRFCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRFCONV_OK 0
#define CRFCONV_ERR_INVALID_ARG (-1)
#define CRFCONV_ERR_WORKSPACE (-2)   /* workspace_bytes smaller than the matching *_workspace_bytes() */
#define CRFCONV_ERR_UNSUPPORTED (-3) /* e.g. K > 32 */
#define CRFCONV_ERR_NO_DEVICE (-4)

int crfconv_abi_version(void);
/* Static string for a status code returned by any function below. */
const char* crfconv_status_string(int status);

/* ------------------------------------------------------------------------------------------------- kNN
 * Replaces utils/nearest_neighbors/knn_.h:4-19 (cpp_knn, cpp_knn_omp, cpp_knn_batch, cpp_knn_batch_omp) and their
 * Cython callers knn.pyx:33-109.  Exact K nearest neighbours, squared L2 in f32 with nanoflann's operation order
 * (nanoflann.hpp:343-346), ascending (distance, index).  dim is fixed to 3.  K <= 32.  If K > npts the trailing
 * slots hold 0 (the observable behaviour of cpp_knn_omp, knn_.cxx:59-67).                                      */

size_t crfconv_knn_workspace_bytes(int64_t batch_size, int64_t npts, int64_t nqueries, int64_t K);

/* Device pointers.  pts [B,npts,3] f32, queries [B,nqueries,3] f32, out_idx [B,nqueries,K] i64. */
int crfconv_knn_batch(const float* pts, int64_t batch_size, int64_t npts, const float* queries, int64_t nqueries,
                      int64_t K, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* HOST pointers — the drop-in for `void cpp_knn_batch(const float* batch_data, size_t batch_size, size_t npts,
 * size_t dim, const float* queries, size_t nqueries, size_t K, long* batch_indices)` (knn_.h:13-15) and, with
 * batch_size = 1, for cpp_knn (knn_.h:4-6): copies to the device, searches, copies back, synchronises. dim must be 3. */
int crfconv_cpp_knn_batch(const float* batch_data, size_t batch_size, size_t npts, size_t dim, const float* queries,
                          size_t nqueries, size_t K, int64_t* batch_indices);

/* ------------------------------------------------------------------------------------- grid subsampling
 * Replaces grid_subsampling() (utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.h:84-91,
 * grid_subsampling.cpp:5-106) and the array handling of wrapper.cpp:58-286.  Barycentre per voxel, mean feature,
 * plurality label; f32 sums in original point order (bit-exact).  order: 0 = ascending voxel key,
 * 1 = the reference's row order (libstdc++ unordered_map iteration order, replayed on the host).
 * Outputs must hold N rows; *M_out receives the number of occupied voxels.  Synchronises (M is data dependent). */

size_t crfconv_grid_subsample_workspace_bytes(int64_t N, int64_t fdim, int64_t ldim);

/* Device pointers.  features / classes / out_keys may be NULL (with fdim / ldim = 0). */
int crfconv_grid_subsample(const float* points, int64_t N, const float* features, int64_t fdim, const int32_t* classes,
                           int64_t ldim, float sampleDl, int order, float* out_points, float* out_features,
                           int32_t* out_classes, unsigned long long* out_keys, int64_t* M_out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* HOST pointers (what wrapper.cpp:205-229 does with its numpy buffers). */
int crfconv_grid_subsample_host(const float* points, int64_t N, const float* features, int64_t fdim,
                                const int32_t* classes, int64_t ldim, float sampleDl, int order, float* out_points,
                                float* out_features, int32_t* out_classes, int64_t* M_out);

#ifdef __cplusplus
}
#endif
#endif /* CRFCONV_B200_H_ */
