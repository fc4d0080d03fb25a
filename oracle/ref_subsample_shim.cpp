// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// extern "C" doorway onto the *unmodified* reference grid-subsampling core
// (/root/reference/utils/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106
// and cpp_utils/cloud/cloud.cpp).  The reference's CPython wrapper (wrapper.cpp) needs the
// numpy-1.x C API and cannot be compiled against numpy 2.x, so the two arithmetic files are
// compiled directly and driven through ctypes.  Two-phase protocol: run() keeps the result in a
// static holder and returns M; fetch() copies it out.
#include <cstring>
#include <vector>
#include "grid_subsampling/grid_subsampling.h"

static std::vector<PointXYZ> g_pts;
static std::vector<float> g_feat;
static std::vector<int> g_cls;

extern "C" {

long ref_grid_subsample_run(const float* points, long N, const float* features, long fdim,
                            const int* classes, long ldim, float dl) {
    std::vector<PointXYZ> op((const PointXYZ*)points, (const PointXYZ*)points + N);
    std::vector<float> of;
    std::vector<int> oc;
    if (features && fdim > 0) of.assign(features, features + N * fdim);
    if (classes && ldim > 0) oc.assign(classes, classes + N * ldim);
    g_pts.clear(); g_feat.clear(); g_cls.clear();
    grid_subsampling(op, g_pts, of, g_feat, oc, g_cls, dl, 0);
    return (long)g_pts.size();
}

void ref_grid_subsample_fetch(float* points, float* features, int* classes) {
    if (points) std::memcpy(points, g_pts.data(), g_pts.size() * sizeof(PointXYZ));
    if (features && !g_feat.empty()) std::memcpy(features, g_feat.data(), g_feat.size() * sizeof(float));
    if (classes && !g_cls.empty()) std::memcpy(classes, g_cls.data(), g_cls.size() * sizeof(int));
}

}  // extern "C"
