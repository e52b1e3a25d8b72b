// Internal (C++) interface between the translation units of the library. The public C ABI is api.cu.
#pragma once
#include "common.cuh"

namespace d3f {

// ---- gemm.cu ----------------------------------------------------------------------------------------
struct Epilogue {
  const float* rowscale;  // [M] or null
  const float* bn_scale;  // [N] or null
  const float* bn_shift;  // [N] or null (used with bn_scale)
  const float* bias;      // [N] or null
  const float* residual;  // [M,N] or null
  float leaky_alpha;      // < 0: none
  const int* row_map;     // [M] or null: GEMM row m is written to output row row_map[m]
  const int* m_dev = nullptr;   // optional: actual row count in device memory (M is then the launch capacity)
  int m_off = 0;                // rows of *m_dev that precede this GEMM's row 0 (KPConv query chunks)
};
int gemm_f32(const float* A, const float* B, float* C, int M, int N, int K, const Epilogue& ep, cudaStream_t stream);

// ---- tc_gemm.cu (tcgen05 / TMEM, 3xTF32) ------------------------------------------------------------
size_t tc_packed_floats(int K, int N);
int tc_padded_n(int N);
int tc_pack_weight(const float* W, int K, int N, float* packed, cudaStream_t stream);
bool tc_gemm_supported(const float* A, int K);
int tc_gemm(const float* A, const float* Bp, float* C, int M, int N, int K, const Epilogue& ep, cudaStream_t stream,
            float* split_ws = nullptr, const float* A2 = nullptr, int K1 = 0);
int tc_gemm_splits(int M, int N, int K);
size_t tc_gemm_split_ws_floats(int M, int N, int K);

// ---- grid.cu ----------------------------------------------------------------------------------------
int launch_batch_start(const int* len, int B, int* start, cudaStream_t stream);
int bbox_device(const float* pts, int N, float* out_bbox, cudaStream_t stream);
size_t grid_subsample_workspace_bytes(int N, int B);
int grid_subsample(const float* pts, const int* batch_len, int B, int N, float dl, const float* feats, int fdim,
                   const int* classes, int ldim, const float* host_bbox, float* out_pts, float* out_feats,
                   int* out_classes, int* out_batch_len, int* out_M, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream, const int* n_dev = nullptr, int out_capacity = -1, int* status = nullptr,
                   const int* start_pre = nullptr);

// ---- neighbors.cu -----------------------------------------------------------------------------------
size_t radius_neighbors_workspace_bytes(int Ns, int B, float radius, const float* host_bbox);
// ns_dev / nq_dev / pad_dev (optional): actual row counts / the shadow index in device memory; Ns / Nq are then capacities
int radius_neighbors_build(const float* supports, const int* s_batch_len, int B, int Ns, float radius,
                           const float* host_bbox, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                           const int* ns_dev = nullptr, const int* s_start_pre = nullptr);
int radius_neighbors_count(const float* queries, const int* q_batch_len, int Nq, int B, int Ns, float radius,
                           const float* host_bbox, const void* workspace, int* counts, int* out_max,
                           cudaStream_t stream);
int radius_neighbors_order(const void* workspace, int Ns, int B, float radius, const float* host_bbox, int* out_order,
                           cudaStream_t stream);
int radius_neighbors_fill(const float* queries, const int* q_batch_len, int Nq, int B, int Ns, float radius,
                          const float* host_bbox, const void* workspace, int cols, int pad_value, int* out_idx,
                          cudaStream_t stream, const int* nq_dev = nullptr, const int* pad_dev = nullptr,
                          const int* q_start_pre = nullptr);

// ---- pyramid.cu -------------------------------------------------------------------------------------
size_t pyramid_workspace_bytes(int B, const d3f_pyramid_spec* spec, const int* capacity, const float* host_bbox);
int pyramid_build(const float* points, const int* lengths, int B, int N0, const d3f_pyramid_spec* spec,
                  const float* host_bbox, float* const* out_points, int* const* out_lengths,
                  int* const* out_neighbors, int* const* out_pools, int* const* out_upsamples, const int* capacity,
                  int* out_level_sizes, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                  int* d_counts = nullptr, int* d_status = nullptr, const int* n0_dev = nullptr);

// ---- kpconv.cu --------------------------------------------------------------------------------------
size_t kpconv_workspace_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout);
int kpconv_forward_impl(bool deform, const float* q, const float* s, const int* idx, const float* feat,
                        const float* Kp, const float* offsets, const float* modulations, const float* W,
                        const float* W_packed, const int* query_order, int Nq,
                        int Ns, int H, int K, int Cin, int Cout, float extent, int influence, int mode, int normalize,
                        const float* bn_scale, const float* bn_shift, const float* bias, float leaky_alpha,
                        float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                        const int* nq_dev = nullptr, const int* ns_dev = nullptr);

// ---- kpconv_fused.cu (one persistent kernel per layer: gather + correlation + tcgen05 contraction) -----
bool kpconv_fused_supported(int Nq, int H, int K, int Cin, int Cout, int influence, int mode, const float* feat,
                            const float* W, const float* out, const int* query_order);
size_t kpconv_fused_workspace_bytes();
int kpconv_fused_forward(const float* q, const float4* s4, const int* idx, const float* feat, const float* Kp,
                         const float* W, float* w_img, int Nq, int Ns, int H, int Cout, float extent, int normalize,
                         const float* bn_scale, const float* bn_shift, const float* bias, float leaky_alpha, float* out,
                         cudaStream_t stream, const int* nq_dev = nullptr, const int* ns_dev = nullptr);

// ---- pool.cu ----------------------------------------------------------------------------------------
int ind_max_pool(const float* x, const int* inds, int N1, int N2, int H, int C, float* out, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream, const int* n1_dev = nullptr, const int* n2_dev = nullptr);
int closest_pool(const float* x, const int* inds, int N1, int N2, int ld_inds, int C, float* out,
                 cudaStream_t stream, const int* n1_dev = nullptr, const int* n2_dev = nullptr);
int l2_normalize(const float* x, int N, int C, float eps, float* out, cudaStream_t stream, const int* n_dev = nullptr);
size_t detection_scores_workspace_bytes(int N, int B);
int detection_scores(const float* feats, const int* neighbors, const int* lengths, int B, int N, int H, int D,
                     float* out_scores, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     const int* n_dev = nullptr);
int affine_leaky(const float* x, int N, int C, const float* scale, const float* shift, const float* residual,
                 float alpha, float* out, cudaStream_t stream, const int* n_dev = nullptr);

}  // namespace d3f
