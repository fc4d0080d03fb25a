// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// extern "C" doorway onto the *unmodified* reference kNN wrapper so that it can be
// driven through ctypes.  The reference sources are compiled where they lie under
// /root/reference/utils/nearest_neighbors (knn_.cxx + nanoflann.hpp v1.2.3); nothing is
// copied into this repository.  The four entry points below forward 1:1 to
// knn_.h:4-27 (cpp_knn, cpp_knn_omp, cpp_knn_batch, cpp_knn_batch_omp, cpp_knn_batch_distance_pick[_omp]).
#include <cstddef>
#include "knn_.h"

extern "C" {

void ref_knn(const float* pts, size_t npts, size_t dim, const float* q, size_t nq, size_t K,
             long* out, int omp) {
    if (omp) cpp_knn_omp(pts, npts, dim, q, nq, K, out);
    else     cpp_knn(pts, npts, dim, q, nq, K, out);
}

void ref_knn_batch(const float* pts, size_t B, size_t npts, size_t dim, const float* q, size_t nq,
                   size_t K, long* out, int omp) {
    if (omp) cpp_knn_batch_omp(pts, B, npts, dim, q, nq, K, out);
    else     cpp_knn_batch(pts, B, npts, dim, q, nq, K, out);
}

// knn_.h:21-27: coverage sampler (seeded with time(0) inside the reference: the picks are not reproducible, the invariants are)
void ref_knn_batch_distance_pick(const float* pts, size_t B, size_t npts, size_t dim, float* queries, size_t nq, size_t K, long* out, int omp) {
    if (omp) cpp_knn_batch_distance_pick_omp(pts, B, npts, dim, queries, nq, K, out);
    else     cpp_knn_batch_distance_pick(pts, B, npts, dim, queries, nq, K, out);
}

}  // extern "C"
