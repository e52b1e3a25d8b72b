// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" shim over the reference's cpp_wrappers grid_subsampling core
// (cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-105), compiled
// from /root/reference by oracle/Makefile. It does what the CPython wrapper does around
// the core (wrapper.cpp:216-265: vectors in, grid_subsampling(...), memcpy out); the
// wrapper itself does not build against NumPy 2.x (NPY_IN_ARRAY, wrapper.cpp:100).
// Built as its own .so because the tf copy defines a different `class SampledData`.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

extern "C" {

void refw_free(void* p) { free(p); }

// features may be NULL (fdim = 0); classes may be NULL (ldim = 0).
// out_* are malloc'ed here; *out_M receives the number of cells.
int refw_grid_subsampling(const float* p, int N, const float* feats, int fdim, const int* classes,
                          int ldim, float dl, int verbose, float** out_pts, float** out_feats,
                          int** out_classes, int* out_M) {
  std::vector<PointXYZ> pts((const PointXYZ*)p, (const PointXYZ*)p + N), sub;
  std::vector<float> f, sf;
  std::vector<int> c, sc;
  if (feats && fdim > 0) f.assign(feats, feats + (size_t)N * fdim);
  if (classes && ldim > 0) c.assign(classes, classes + (size_t)N * ldim);
  grid_subsampling(pts, sub, f, sf, c, sc, dl, verbose);
  int M = (int)sub.size();
  *out_M = M;
  *out_pts = (float*)malloc(sizeof(float) * (3 * (size_t)M + 1));
  memcpy(*out_pts, sub.data(), sizeof(float) * 3 * (size_t)M);
  *out_feats = NULL;
  *out_classes = NULL;
  if (!f.empty()) {
    *out_feats = (float*)malloc(sizeof(float) * (sf.size() + 1));
    memcpy(*out_feats, sf.data(), sizeof(float) * sf.size());
  }
  if (!c.empty()) {
    *out_classes = (int*)malloc(sizeof(int) * (sc.size() + 1));
    memcpy(*out_classes, sc.data(), sizeof(int) * sc.size());
  }
  return 0;
}

}  // extern "C"
