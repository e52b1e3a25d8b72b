// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" shim over the reference's own CPU cores, compiled *from the sources
// where they lie under /root/reference* (see oracle/Makefile). Nothing from the
// reference is copied into this repository; this file only adapts flat C arrays
// to the std::vector<PointXYZ> signatures, exactly like the TF op shells do:
//   tf_custom_ops/tf_neighbors/tf_batch_neighbors.cpp:76-85   (copy-in)
//   tf_custom_ops/tf_neighbors/tf_batch_neighbors.cpp:109-115 (copy-out)
//   tf_custom_ops/tf_subsampling/tf_batch_subsampling.cpp:57-72, 99-121
//   tf_custom_ops/tf_neighbors/tf_neighbors.cpp:40-86
//   tf_custom_ops/tf_subsampling/tf_subsampling.cpp:40-86
//
// Entry points (all return a malloc'ed buffer the caller frees with ref_free):
//   ref_batch_neighbors   -> batch_nanoflann_neighbors (neighbors.cpp:211-332)
//   ref_ordered_neighbors -> ordered_neighbors         (neighbors.cpp:58-123)
//   ref_batch_subsampling -> batch_grid_subsampling    (grid_subsampling.cpp:101-149, tf copy)
//   ref_grid_subsampling  -> grid_subsampling          (grid_subsampling.cpp:5-97,   tf copy)
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tf_neighbors/neighbors/neighbors.h"
#include "tf_subsampling/grid_subsampling/grid_subsampling.h"

static std::vector<PointXYZ> to_points(const float* p, int n) {
  std::vector<PointXYZ> v;
  v.reserve(n);
  for (int i = 0; i < n; ++i) v.push_back(PointXYZ(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
  return v;
}

extern "C" {

void ref_free(void* p) { free(p); }

// returns int32[Nq * (*out_cols)]
int* ref_batch_neighbors(const float* q, int Nq, const float* s, int Ns, const int* qb,
                         const int* sb, int B, float radius, int* out_cols) {
  std::vector<PointXYZ> queries = to_points(q, Nq), supports = to_points(s, Ns);
  std::vector<int> q_batches(qb, qb + B), s_batches(sb, sb + B);
  std::vector<int> idx;
  batch_nanoflann_neighbors(queries, supports, q_batches, s_batches, idx, radius);
  int cols = Nq > 0 ? (int)(idx.size() / (size_t)Nq) : 0;
  *out_cols = cols;
  int* out = (int*)malloc(sizeof(int) * (idx.size() + 1));
  memcpy(out, idx.data(), sizeof(int) * idx.size());
  return out;
}

int* ref_ordered_neighbors(const float* q, int Nq, const float* s, int Ns, float radius,
                           int* out_cols) {
  std::vector<PointXYZ> queries = to_points(q, Nq), supports = to_points(s, Ns);
  std::vector<int> idx;
  ordered_neighbors(queries, supports, idx, radius);
  int cols = Nq > 0 ? (int)(idx.size() / (size_t)Nq) : 0;
  *out_cols = cols;
  int* out = (int*)malloc(sizeof(int) * (idx.size() + 1));
  memcpy(out, idx.data(), sizeof(int) * idx.size());
  return out;
}

// returns float[3 * (*out_M)]; out_batches[B] receives the new stack lengths
float* ref_batch_subsampling(const float* p, int N, const int* batches, int B, float dl,
                             int* out_batches, int* out_M) {
  std::vector<PointXYZ> pts = to_points(p, N), sub;
  std::vector<float> f, sf;
  std::vector<int> c, sc;
  std::vector<int> ob(batches, batches + B), nb;
  batch_grid_subsampling(pts, sub, f, sf, c, sc, ob, nb, dl);
  for (int b = 0; b < B; ++b) out_batches[b] = nb[b];
  *out_M = (int)sub.size();
  float* out = (float*)malloc(sizeof(float) * (3 * sub.size() + 1));
  for (size_t i = 0; i < sub.size(); ++i) {
    out[3 * i] = sub[i].x;
    out[3 * i + 1] = sub[i].y;
    out[3 * i + 2] = sub[i].z;
  }
  return out;
}

float* ref_grid_subsampling(const float* p, int N, float dl, int* out_M) {
  std::vector<PointXYZ> pts = to_points(p, N), sub;
  std::vector<float> f, sf;
  std::vector<int> c, sc;
  grid_subsampling(pts, sub, f, sf, c, sc, dl);
  *out_M = (int)sub.size();
  float* out = (float*)malloc(sizeof(float) * (3 * sub.size() + 1));
  for (size_t i = 0; i < sub.size(); ++i) {
    out[3 * i] = sub[i].x;
    out[3 * i + 1] = sub[i].y;
    out[3 * i + 2] = sub[i].z;
  }
  return out;
}

}  // extern "C"
