// extern "C" entry points of libd3feat_b200.so (declared in include/d3feat_b200.h).
#include <stdarg.h>

#include "ops.cuh"

namespace d3f {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return D3F_ERR_CUDA;
}

void count_launch(int n) { g_launches += n; }

}  // namespace d3f

using namespace d3f;

extern "C" {

int d3f_version(void) { return 100; }

const char* d3f_last_error(void) { return g_err; }

long long d3f_launch_count(void) { return g_launches; }

int d3f_bbox(const float* pts, int N, float* out_bbox, d3f_stream_t stream) {
  D3F_REQUIRE(pts != nullptr || N == 0, D3F_ERR_INVALID, "d3f_bbox: null points");
  D3F_REQUIRE(out_bbox != nullptr && N >= 0, D3F_ERR_INVALID, "d3f_bbox: bad arguments");
  return bbox_device(pts, N, out_bbox, (cudaStream_t)stream);
}

size_t d3f_grid_subsample_workspace_bytes(int N, int B) { return grid_subsample_workspace_bytes(N, B); }

int d3f_grid_subsample(const float* pts, const int* batch_len, int B, int N, float dl, const float* feats, int fdim,
                       const int* classes, int ldim, const float* host_bbox, float* out_pts, float* out_feats,
                       int* out_classes, int* out_batch_len, int* out_M, void* workspace, size_t workspace_bytes,
                       d3f_stream_t stream) {
  D3F_REQUIRE((pts != nullptr || N == 0) && batch_len != nullptr && out_pts != nullptr && out_batch_len != nullptr &&
                  out_M != nullptr && workspace != nullptr,
              D3F_ERR_INVALID, "d3f_grid_subsample: null pointer");
  D3F_REQUIRE((fdim == 0 || out_feats != nullptr) && (ldim == 0 || out_classes != nullptr), D3F_ERR_INVALID,
              "d3f_grid_subsample: missing feature / class output");
  return grid_subsample(pts, batch_len, B, N, dl, feats, fdim, classes, ldim, host_bbox, out_pts, out_feats,
                        out_classes, out_batch_len, out_M, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t d3f_radius_neighbors_workspace_bytes(int Ns, int B, float radius, const float* host_bbox) {
  return radius_neighbors_workspace_bytes(Ns, B, radius, host_bbox);
}

int d3f_radius_neighbors_build(const float* supports, const int* s_batch_len, int B, int Ns, float radius,
                               const float* host_bbox, void* workspace, size_t workspace_bytes, d3f_stream_t stream) {
  D3F_REQUIRE((supports != nullptr || Ns == 0) && s_batch_len != nullptr && workspace != nullptr, D3F_ERR_INVALID,
              "d3f_radius_neighbors_build: null pointer");
  return radius_neighbors_build(supports, s_batch_len, B, Ns, radius, host_bbox, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

int d3f_radius_neighbors_count(const float* queries, const int* q_batch_len, int Nq, const float* supports,
                               const int* s_batch_len, int B, int Ns, float radius, const float* host_bbox,
                               const void* workspace, int* counts, int* out_max, d3f_stream_t stream) {
  (void)supports;
  (void)s_batch_len;
  D3F_REQUIRE((queries != nullptr || Nq == 0) && q_batch_len != nullptr && workspace != nullptr, D3F_ERR_INVALID,
              "d3f_radius_neighbors_count: null pointer");
  return radius_neighbors_count(queries, q_batch_len, Nq, B, Ns, radius, host_bbox, workspace, counts, out_max,
                                (cudaStream_t)stream);
}

int d3f_radius_neighbors_fill(const float* queries, const int* q_batch_len, int Nq, const float* supports,
                              const int* s_batch_len, int B, int Ns, float radius, const float* host_bbox,
                              const void* workspace, int cols, int pad_value, int* out_idx, d3f_stream_t stream) {
  (void)supports;
  (void)s_batch_len;
  D3F_REQUIRE((queries != nullptr || Nq == 0) && q_batch_len != nullptr && workspace != nullptr, D3F_ERR_INVALID,
              "d3f_radius_neighbors_fill: null pointer");
  return radius_neighbors_fill(queries, q_batch_len, Nq, B, Ns, radius, host_bbox, workspace, cols, pad_value, out_idx,
                               (cudaStream_t)stream);
}

int d3f_radius_neighbors_order(const void* workspace, int Ns, int B, float radius, const float* host_bbox,
                               int* out_order, d3f_stream_t stream) {
  D3F_REQUIRE(workspace != nullptr && (out_order != nullptr || Ns == 0), D3F_ERR_INVALID,
              "d3f_radius_neighbors_order: null pointer");
  return radius_neighbors_order(workspace, Ns, B, radius, host_bbox, out_order, (cudaStream_t)stream);
}

size_t d3f_kpconv_workspace_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout) {
  return kpconv_workspace_bytes(Nq, Ns, H, K, Cin, Cout);
}

size_t d3f_pyramid_workspace_bytes(int B, const d3f_pyramid_spec* spec, const int* capacity, const float* host_bbox) {
  return pyramid_workspace_bytes(B, spec, capacity, host_bbox);
}

int d3f_pyramid_build(const float* points, const int* lengths, int B, int N0, const d3f_pyramid_spec* spec,
                      const float* host_bbox, float* const* out_points, int* const* out_lengths,
                      int* const* out_neighbors, int* const* out_pools, int* const* out_upsamples, const int* capacity,
                      int* out_level_sizes, void* workspace, size_t workspace_bytes, d3f_stream_t stream,
                      int* d_counts, int* d_status, const int* n0_dev) {
  D3F_REQUIRE((points != nullptr || N0 == 0) && lengths != nullptr && out_points && out_lengths && out_neighbors &&
                  out_pools && out_upsamples && workspace,
              D3F_ERR_INVALID, "d3f_pyramid_build: null pointer");
  D3F_REQUIRE(B >= 1 && B <= kMaxBatch, D3F_ERR_INVALID, "d3f_pyramid_build: B=%d", B);
  return pyramid_build(points, lengths, B, N0, spec, host_bbox, out_points, out_lengths, out_neighbors, out_pools,
                       out_upsamples, capacity, out_level_sizes, workspace, workspace_bytes, (cudaStream_t)stream,
                       d_counts, d_status, n0_dev);
}

size_t d3f_packed_weight_floats(int K, int N) { return tc_packed_floats(K, N); }

int d3f_pack_weight(const float* W, int K, int N, float* packed, d3f_stream_t stream) {
  return tc_pack_weight(W, K, N, packed, (cudaStream_t)stream);
}

int d3f_kpconv_forward(const float* q, const float* s, const int* idx, const float* feat, const float* Kp,
                       const float* W, const float* W_packed, const int* query_order, int Nq, int Ns, int H, int K, int Cin, int Cout, float extent, int influence,
                       int mode, int normalize, const float* bn_scale, const float* bn_shift, const float* bias,
                       float leaky_alpha, float* out, void* workspace, size_t workspace_bytes, d3f_stream_t stream,
                       const int* nq_dev, const int* ns_dev) {
  D3F_REQUIRE(Nq == 0 || (q && s && idx && feat && Kp && W && out && workspace), D3F_ERR_INVALID,
              "d3f_kpconv_forward: null pointer");
  return kpconv_forward_impl(false, q, s, idx, feat, Kp, nullptr, nullptr, W, W_packed, query_order, Nq, Ns, H, K, Cin, Cout, extent,
                             influence, mode, normalize, bn_scale, bn_shift, bias, leaky_alpha, out, workspace,
                             workspace_bytes, (cudaStream_t)stream, nq_dev, ns_dev);
}

int d3f_kpconv_deform_forward(const float* q, const float* s, const int* idx, const float* feat, const float* Kp,
                              const float* offsets, const float* modulations, const float* W, const float* W_packed,
                              const int* query_order, int Nq, int Ns, int H, int K, int Cin, int Cout, float extent, int influence, int mode, const float* bn_scale,
                              const float* bn_shift, const float* bias, float leaky_alpha, float* out,
                              void* workspace, size_t workspace_bytes, d3f_stream_t stream, const int* nq_dev,
                              const int* ns_dev) {
  D3F_REQUIRE(Nq == 0 || (q && s && idx && feat && Kp && W && out && workspace && offsets), D3F_ERR_INVALID,
              "d3f_kpconv_deform_forward: null pointer");
  return kpconv_forward_impl(true, q, s, idx, feat, Kp, offsets, modulations, W, W_packed, query_order, Nq, Ns, H, K, Cin, Cout, extent,
                             influence, mode, 0, bn_scale, bn_shift, bias, leaky_alpha, out, workspace,
                             workspace_bytes, (cudaStream_t)stream, nq_dev, ns_dev);
}

int d3f_unary_forward(const float* x, const float* W, const float* W_packed, int N, int Cin, int Cout,
                      const float* bn_scale,
                      const float* bn_shift, const float* bias, const float* residual, float leaky_alpha, float* out,
                      d3f_stream_t stream, const int* n_dev) {
  D3F_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1, D3F_ERR_INVALID, "d3f_unary_forward: bad shape N=%d Cin=%d Cout=%d", N,
              Cin, Cout);
  D3F_REQUIRE(N == 0 || (x && W && out), D3F_ERR_INVALID, "d3f_unary_forward: null pointer");
  D3F_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), D3F_ERR_INVALID,
              "d3f_unary_forward: bn_scale/bn_shift mismatch");
  Epilogue ep;
  ep.rowscale = nullptr;
  ep.bn_scale = bn_scale;
  ep.bn_shift = bn_shift;
  ep.bias = bias;
  ep.residual = residual;
  ep.leaky_alpha = leaky_alpha;
  ep.row_map = nullptr;
  ep.m_dev = n_dev;
  if (W_packed != nullptr && tc_gemm_supported(x, Cin))
    return tc_gemm(x, W_packed, out, N, Cout, Cin, ep, (cudaStream_t)stream);
  return gemm_f32(x, W, out, N, Cout, Cin, ep, (cudaStream_t)stream);
}

int d3f_unary_pair_forward(const float* x1, int Cin1, const float* x2, int Cin2, const float* W_packed, int N,
                           int Cout, const float* shift, float leaky_alpha, float* out, d3f_stream_t stream,
                           const int* n_dev) {
  D3F_REQUIRE(N >= 0 && Cin1 >= 1 && Cin2 >= 1 && Cout >= 1, D3F_ERR_INVALID,
              "d3f_unary_pair_forward: bad shape N=%d Cin=%d+%d Cout=%d", N, Cin1, Cin2, Cout);
  D3F_REQUIRE(N == 0 || (x1 && x2 && W_packed && out), D3F_ERR_INVALID, "d3f_unary_pair_forward: null pointer");
  D3F_REQUIRE(Cin1 % 32 == 0 && Cin2 % 4 == 0, D3F_ERR_INVALID,
              "d3f_unary_pair_forward: Cin1 must be a multiple of 32 and Cin2 of 4 (got %d, %d)", Cin1, Cin2);
  Epilogue ep;
  ep.rowscale = nullptr;
  ep.bn_scale = nullptr;
  ep.bn_shift = nullptr;
  ep.bias = shift;
  ep.residual = nullptr;
  ep.leaky_alpha = leaky_alpha;
  ep.row_map = nullptr;
  ep.m_dev = n_dev;
  return tc_gemm(x1, W_packed, out, N, Cout, Cin1 + Cin2, ep, (cudaStream_t)stream, nullptr, x2, Cin1);
}

size_t d3f_ind_max_pool_workspace_bytes(int C) { return sizeof(unsigned) * ((size_t)(C > 0 ? C : 1) + 1); }

int d3f_ind_max_pool(const float* x, const int* inds, int N1, int N2, int H, int C, float* out, void* workspace,
                     size_t workspace_bytes, d3f_stream_t stream, const int* n1_dev, const int* n2_dev) {
  D3F_REQUIRE(N2 == 0 || (x && inds && out && workspace), D3F_ERR_INVALID, "d3f_ind_max_pool: null pointer");
  return ind_max_pool(x, inds, N1, N2, H, C, out, workspace, workspace_bytes, (cudaStream_t)stream, n1_dev, n2_dev);
}

int d3f_closest_pool(const float* x, const int* inds, int N1, int N2, int ld_inds, int C, float* out,
                     d3f_stream_t stream, const int* n1_dev, const int* n2_dev) {
  D3F_REQUIRE(N2 == 0 || (x && inds && out), D3F_ERR_INVALID, "d3f_closest_pool: null pointer");
  return closest_pool(x, inds, N1, N2, ld_inds, C, out, (cudaStream_t)stream, n1_dev, n2_dev);
}

int d3f_l2_normalize(const float* x, int N, int C, float eps, float* out, d3f_stream_t stream, const int* n_dev) {
  D3F_REQUIRE(N == 0 || (x && out), D3F_ERR_INVALID, "d3f_l2_normalize: null pointer");
  return l2_normalize(x, N, C, eps, out, (cudaStream_t)stream, n_dev);
}

size_t d3f_detection_scores_workspace_bytes(int N, int B) { return detection_scores_workspace_bytes(N, B); }

int d3f_detection_scores(const float* feats, const int* neighbors, const int* lengths, int B, int N, int H, int D,
                         float* out_scores, void* workspace, size_t workspace_bytes, d3f_stream_t stream,
                         const int* n_dev) {
  D3F_REQUIRE(N == 0 || (feats && neighbors && lengths && out_scores && workspace), D3F_ERR_INVALID,
              "d3f_detection_scores: null pointer");
  return detection_scores(feats, neighbors, lengths, B, N, H, D, out_scores, workspace, workspace_bytes,
                          (cudaStream_t)stream, n_dev);
}

int d3f_affine_leaky(const float* x, int N, int C, const float* scale, const float* shift, const float* residual,
                     float leaky_alpha, float* out, d3f_stream_t stream, const int* n_dev) {
  D3F_REQUIRE(N == 0 || (x && out), D3F_ERR_INVALID, "d3f_affine_leaky: null pointer");
  return affine_leaky(x, N, C, scale, shift, residual, leaky_alpha, out, (cudaStream_t)stream, n_dev);
}

}  // extern "C"
