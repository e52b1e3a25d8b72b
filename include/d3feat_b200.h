/* d3feat_b200 -- C ABI of the Blackwell-native (sm_100a) D3Feat hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / TF types. Every entry point
 *   - takes DEVICE pointers unless a parameter is explicitly marked "host",
 *   - takes an explicit CUDA stream (cudaStream_t passed as void*), enqueues its work there and
 *     returns without synchronising unless stated otherwise,
 *   - never frees or allocates caller-visible memory: scratch comes from the caller-supplied
 *     workspace (size from the matching *_workspace_bytes query),
 *   - returns 0 (D3F_OK) or a negative D3F_ERR_* code; d3f_last_error() gives the message
 *     (thread-local). No exception crosses the boundary.
 *   - keeps no global mutable state: re-entrant across host threads and streams (the reference's
 *     OpKernels are invoked concurrently by the tf.data map pool, datasets/common.py:600,744).
 *
 * Reference interfaces replaced (paths relative to the D3Feat repository):
 *   d3f_grid_subsample        tf_custom_ops/tf_subsampling/tf_batch_subsampling.cpp:8-20,30-122
 *                             tf_custom_ops/tf_subsampling/tf_subsampling.cpp:8-17 (B = 1)
 *                             cpp_wrappers/cpp_subsampling/wrapper.cpp:58-286 (features / classes)
 *   d3f_radius_neighbors_*    tf_custom_ops/tf_neighbors/tf_batch_neighbors.cpp:8-30,40-116
 *                             tf_custom_ops/tf_neighbors/tf_neighbors.cpp:8-18 (B = 1, pad = -1)
 *   d3f_kpconv_forward        kernels/convolution_ops.py:161-255 (KPConv_ops) + BN/LeakyReLU epilogue
 *                             models/network_blocks.py:149-165,185-186
 *   d3f_kpconv_deform_forward kernels/convolution_ops.py:379-499 (KPConv_deform_ops)
 *   d3f_unary_forward         kernels/convolution_ops.py:90-99 + models/network_blocks.py:207-219,
 *                             :343-368 (conv3 + shortcut add + LeakyReLU)
 *   d3f_ind_max_pool          models/network_blocks.py:51-66
 *   d3f_closest_pool          models/network_blocks.py:69-83
 *   d3f_l2_normalize          models/D3Feat.py:65
 */
#ifndef D3FEAT_B200_H_
#define D3FEAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3F_OK 0
#define D3F_ERR_INVALID (-1)   /* bad argument (shape, enum, null pointer)            */
#define D3F_ERR_CAPACITY (-2)  /* caller-supplied output capacity / grid budget too small */
#define D3F_ERR_CUDA (-3)      /* CUDA runtime error (message in d3f_last_error)       */
#define D3F_ERR_WORKSPACE (-4) /* workspace smaller than *_workspace_bytes             */

/* KP_influence / aggregation_mode enums (kernels/convolution_ops.py:208-232) */
#define D3F_INFLUENCE_CONSTANT 0
#define D3F_INFLUENCE_LINEAR 1
#define D3F_INFLUENCE_GAUSSIAN 2
#define D3F_MODE_SUM 0
#define D3F_MODE_CLOSEST 1

typedef void* d3f_stream_t; /* cudaStream_t */

int d3f_version(void);
const char* d3f_last_error(void);
/* number of kernels this library has launched from the calling host thread (bench.py "gpu_launches") */
long long d3f_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Bounding box of a stacked cloud: out_bbox[6] = {minx,miny,minz,maxx,maxy,maxz} (device floats).
 * Used to bound the hash grids of the two ops below without a host round trip per call.
 * ------------------------------------------------------------------------------------------- */
int d3f_bbox(const float* pts, int N, float* out_bbox, d3f_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Grid subsampling (voxel barycenters), stacked clouds.
 *   pts[N,3], batch_len[B] (device int32), dl: cell size.
 *   feats[N,fdim] / classes[N,ldim] optional (NULL, 0): the cpp_wrappers signature.
 *   host_bbox: host float[6] bounding ALL points (from d3f_bbox or known a priori). It only bounds
 *     the sort-key width; the per-cloud origin is recomputed exactly on the device.
 *   Outputs (capacity N rows each): out_pts[<=N,3], out_feats, out_classes, out_batch_len[B],
 *     out_M[1] (device int32: total number of cells). Cells are emitted per cloud in ascending
 *     reference cell key iX + NX*iY + NX*NY*iZ (grid_subsampling.cpp:53-56); barycenters are
 *     bit-identical to the reference (fp32 sums in input order, * (float)(1.0/count)).
 *   classes follow the reference literally: the LARGEST label present in the cell
 *     (std::max_element over map pairs, grid_subsampling.cpp:97-101).
 * ------------------------------------------------------------------------------------------- */
size_t d3f_grid_subsample_workspace_bytes(int N, int B);
int d3f_grid_subsample(const float* pts, const int* batch_len, int B, int N, float dl,
                       const float* feats, int fdim, const int* classes, int ldim,
                       const float* host_bbox, float* out_pts, float* out_feats, int* out_classes,
                       int* out_batch_len, int* out_M, void* workspace, size_t workspace_bytes,
                       d3f_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Radius neighbours, stacked clouds (hash grid over the supports, 27-cell scan per query).
 *   Result-set semantics of the reference (neighbors.cpp:211-332 / nanoflann.hpp:249-253,432-440):
 *   supports of the same cloud with d2 < radius*radius, d2 = ((dx*dx)+dy*dy)+dz*dz in fp32 without
 *   FMA contraction, rows sorted by ascending (d2, index), global support indices, rows padded with
 *   pad_value (Ns for the batch op, -1 for the non-batch op).
 *
 *   Two-phase use for the exact reference shape [Nq, max count]:
 *     d3f_radius_neighbors_build  -> grid over the supports in `workspace`
 *     d3f_radius_neighbors_count  -> counts[Nq] and out_max[1] (device); caller reads out_max
 *     d3f_radius_neighbors_fill   -> out_idx[Nq, cols]; rows longer than cols keep the nearest cols
 *   Single-phase use with a known column cap (the pyramid's neighborhood_limits,
 *   datasets/common.py:399-406): build + fill with cols = cap (count optional).
 * ------------------------------------------------------------------------------------------- */
size_t d3f_radius_neighbors_workspace_bytes(int Ns, int B, float radius, const float* host_bbox);
int d3f_radius_neighbors_build(const float* supports, const int* s_batch_len, int B, int Ns,
                               float radius, const float* host_bbox, void* workspace,
                               size_t workspace_bytes, d3f_stream_t stream);
int d3f_radius_neighbors_count(const float* queries, const int* q_batch_len, int Nq,
                               const float* supports, const int* s_batch_len, int B, int Ns,
                               float radius, const float* host_bbox, const void* workspace,
                               int* counts, int* out_max, d3f_stream_t stream);
/* out_order[Ns]: the support indices in hash-grid cell order (a spatially coherent visiting order). Passing it
 * as `query_order` to d3f_kpconv_forward when queries == supports makes neighbouring queries share their
 * gathered rows in L1/L2; results are unchanged (each query still writes its own output row). */
int d3f_radius_neighbors_order(const void* workspace, int Ns, int B, float radius,
                               const float* host_bbox, int* out_order, d3f_stream_t stream);
int d3f_radius_neighbors_fill(const float* queries, const int* q_batch_len, int Nq,
                              const float* supports, const int* s_batch_len, int B, int Ns,
                              float radius, const float* host_bbox, const void* workspace, int cols,
                              int pad_value, int* out_idx, d3f_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The whole input pyramid of the encoder in ONE call: the loop of Dataset.tf_descriptor_input
 * (datasets/common.py:1325-1397) with the neighbourhood caps of big_neighborhood_filter (:399-406).
 * Level l: neighbors[l] = search(points_l, points_l, conv_radius[l]) (skipped if conv_radius <= 0);
 * if sub_dl[l] > 0: points_{l+1} = grid_subsample(points_l, sub_dl[l]); pools[l] = search(points_{l+1}, points_l,
 * pool_radius[l]); upsamples[l] = search(points_l, points_{l+1}, up_radius[l]). Every index matrix has exactly
 * limit[l] columns (nearest first, padded with the number of supports). Output buffers are caller-allocated with
 * `capacity[l]` rows per level.
 *   Exact form (out_level_sizes != NULL, a HOST int[n_levels]): receives the actual row counts; the call synchronises
 *   the stream once per subsampled level (the next level's launch sizes depend on the cell count) -- what the TF ops
 *   do (their output shapes are data dependent, tf_batch_subsampling.cpp:99-104).
 *   Static form (out_level_sizes == NULL): no device->host read at all. Launches are sized by capacity[l], every
 *   kernel reads its row count from d_counts[l] (DEVICE int[n_levels], written by this call), conditions that the exact
 *   form reports as errors are OR-ed into *d_status (DEVICE int: bit 0 = points outside host_bbox, bit 1 = a level
 *   has more cells than capacity[l+1]). The launch sequence depends only on (B, capacity, spec, host_bbox), so a
 *   caller may capture it in a CUDA graph and replay it for every batch of the same bucket. d_counts[0] = N0, or
 *   *n0_dev when n0_dev != NULL (the level-0 count kept on the device; N0 is then the capacity of `points`).
 *   d_counts / d_status may be NULL in the exact form.
 * ------------------------------------------------------------------------------------------- */
#define D3F_MAX_LEVELS 8
typedef struct {
  int n_levels;
  float conv_radius[D3F_MAX_LEVELS];
  float sub_dl[D3F_MAX_LEVELS];
  float pool_radius[D3F_MAX_LEVELS];
  float up_radius[D3F_MAX_LEVELS];
  int limit[D3F_MAX_LEVELS];
} d3f_pyramid_spec;
size_t d3f_pyramid_workspace_bytes(int B, const d3f_pyramid_spec* spec, const int* capacity,
                                   const float* host_bbox);
int d3f_pyramid_build(const float* points, const int* lengths, int B, int N0,
                      const d3f_pyramid_spec* spec, const float* host_bbox, float* const* out_points,
                      int* const* out_lengths, int* const* out_neighbors, int* const* out_pools,
                      int* const* out_upsamples, const int* capacity, int* out_level_sizes,
                      void* workspace, size_t workspace_bytes, d3f_stream_t stream, int* d_counts,
                      int* d_status, const int* n0_dev);

/* ---------------------------------------------------------------------------------------------
 * Device-side row counts. Every row-wise entry point below takes trailing `const int* ..._dev` arguments (all may
 * be NULL). When given, the integer row count (Nq / Ns / N / N1 / N2) is the CAPACITY of the buffers and of the launch,
 * and the kernels read the actual count from device memory (e.g. &d_counts[l] of d3f_pyramid_build) -- rows beyond
 * it are neither read nor written, and the shadow index is the actual count. This is what lets a whole step be
 * enqueued (or graph-replayed) without the host ever knowing the level sizes.
 * ------------------------------------------------------------------------------------------- */

/* ---------------------------------------------------------------------------------------------
 * Static weights for the tensor-core path. A weight matrix W[K,N] (row-major; for KPConv the [K*Cin, Cout]
 * view of K_values[K,Cin,Cout]) is packed ONCE into the K-major TF32 hi/lo images the tcgen05 kernels
 * consume (3xTF32 split: fp32-level accuracy on the 5th-gen tensor cores). Every forward entry point takes
 * the packed image as an optional `W_packed` argument: NULL selects the CUDA-core fp32 path (same results
 * within rounding), non-NULL the tcgen05 path.
 * ------------------------------------------------------------------------------------------- */
size_t d3f_packed_weight_floats(int K, int N);
int d3f_pack_weight(const float* W, int K, int N, float* packed, d3f_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Rigid KPConv forward (KPConv_ops, convolution_ops.py:161-255), fused with the block epilogue.
 *   q[Nq,3], s[Ns,3], idx[Nq,H] (shadow index = Ns), feat[Ns,Cin], Kp[K,3], W[K,Cin,Cout].
 *   out[Nq,Cout] = epilogue( (sum_k (sum_h w[n,h,k] feat[idx[n,h]]) W_k) / nn[n] )
 *   nn = max(#neighbours whose feature-row sum > 0, 1)   (normalize != 0; :249-253)
 *   epilogue: y = x*bn_scale[c] + bn_shift[c] (if bn_scale != NULL; inference batch norm folded by
 *   the caller: scale = gamma/sqrt(var+1e-6), shift = beta - mean*scale), then + bias[c] (if
 *   bias != NULL), then LeakyReLU(leaky_alpha) if leaky_alpha >= 0 (pass -1 for none).
 *   shadow_xyz: coordinate of the shadow support (1e6 rigid :190, 1000 deformable :414).
 *   K = num_kernel_points (utils/config.py): any value in [1, 64]; K = 15 (the D3Feat configuration) runs the
 *   specialised tensor-core kernels, other values a generic CUDA-core stage 1.
 * ------------------------------------------------------------------------------------------- */
size_t d3f_kpconv_workspace_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout);
int d3f_kpconv_forward(const float* q, const float* s, const int* idx, const float* feat,
                       const float* Kp, const float* W, const float* W_packed, const int* query_order,
                       int Nq, int Ns, int H, int K, int Cin,
                       int Cout, float extent, int influence, int mode, int normalize,
                       const float* bn_scale, const float* bn_shift, const float* bias,
                       float leaky_alpha, float* out, void* workspace, size_t workspace_bytes,
                       d3f_stream_t stream, const int* nq_dev, const int* ns_dev);

/* Deformable KPConv second stage (KPConv_deform_ops, :379-499): per-query kernel points
 * Kp + offsets[n,K,3]; influence distance /extent (no factor 2); neighbours in range of no kernel
 * point are dropped (:435-451); optional modulations[n,K]; no neighbour-count normalisation. */
int d3f_kpconv_deform_forward(const float* q, const float* s, const int* idx, const float* feat,
                              const float* Kp, const float* offsets, const float* modulations,
                              const float* W, const float* W_packed, const int* query_order, int Nq,
                              int Ns, int H, int K, int Cin, int Cout,
                              float extent, int influence, int mode, const float* bn_scale,
                              const float* bn_shift, const float* bias, float leaky_alpha,
                              float* out, void* workspace, size_t workspace_bytes,
                              d3f_stream_t stream, const int* nq_dev, const int* ns_dev);

/* ---------------------------------------------------------------------------------------------
 * Unary convolution (features @ W) with fused epilogue:
 *   y = x@W; y = y*bn_scale + bn_shift (opt); y += bias (opt); y += residual[N,Cout] (opt);
 *   y = LeakyReLU(y) if leaky_alpha >= 0.
 * ------------------------------------------------------------------------------------------- */
int d3f_unary_forward(const float* x, const float* W, const float* W_packed, int N, int Cin, int Cout,
                      const float* bn_scale, const float* bn_shift, const float* bias,
                      const float* residual, float leaky_alpha, float* out, d3f_stream_t stream,
                      const int* n_dev);

/* out[N2,C] = max_h x'[inds[n,h]] with x' = x || colmin(x) (shadow index = N1).
 * workspace: C floats. */
size_t d3f_ind_max_pool_workspace_bytes(int C);
int d3f_ind_max_pool(const float* x, const int* inds, int N1, int N2, int H, int C, float* out,
                     void* workspace, size_t workspace_bytes, d3f_stream_t stream, const int* n1_dev,
                     const int* n2_dev);

/* out[N2,C] = x'[inds[n,0]] with x' = x || zeros. `ld_inds` = row stride of inds (H). */
int d3f_closest_pool(const float* x, const int* inds, int N1, int N2, int ld_inds, int C,
                     float* out, d3f_stream_t stream, const int* n1_dev, const int* n2_dev);

/* out[n,:] = x[n,:] * rsqrt(max(sum x^2, eps)) (tf.nn.l2_normalize, models/D3Feat.py:65) */
int d3f_l2_normalize(const float* x, int N, int C, float eps, float* out, d3f_stream_t stream,
                     const int* n_dev);

/* Two unary convolutions that are summed -- the tail of every resnetb block (network_blocks.py:343-368: conv3 + BN,
 * shortcut unary + BN, add, LeakyReLU) -- as ONE tensor-core GEMM over the concatenated K:
 *   out = leaky([x1 | x2] @ W + shift),  W_packed = d3f_pack_weight of the [Cin1 + Cin2, Cout] matrix whose rows
 * are the two weight matrices with their batch-norm scales folded in, shift = the sum of the two BN shifts.
 * Neither the shortcut tensor nor the [x1 | x2] concatenation is ever materialised. Cin1 % 32 == 0, Cin2 % 4 == 0. */
int d3f_unary_pair_forward(const float* x1, int Cin1, const float* x2, int Cin2, const float* W_packed, int N,
                           int Cout, const float* shift, float leaky_alpha, float* out, d3f_stream_t stream,
                           const int* n_dev);

/* Detection score of D3Feat (models/D3Feat.py:67-115) for B stacked clouds: feats[N,D] are the decoder outputs BEFORE
 * l2 normalisation, neighbors[N,H] the level-0 conv neighbours (shadow index = N), lengths[B] the stack lengths.
 * out_scores[N]. The reference hard-codes B = 2 (anchor || positive); the result is identical for B = 2. */
size_t d3f_detection_scores_workspace_bytes(int N, int B);
int d3f_detection_scores(const float* feats, const int* neighbors, const int* lengths, int B, int N,
                         int H, int D, float* out_scores, void* workspace, size_t workspace_bytes,
                         d3f_stream_t stream, const int* n_dev);

/* Stand-alone block epilogue for callers that do not use the fused forms
 * (models/network_blocks.py:149-165 batch_norm inference form, :185-186 leaky_relu, :368 residual add):
 * y = x*scale[c] + shift[c] (if scale) ; y += residual (if) ; LeakyReLU(leaky_alpha) if >= 0. */
int d3f_affine_leaky(const float* x, int N, int C, const float* scale, const float* shift,
                     const float* residual, float leaky_alpha, float* out, d3f_stream_t stream,
                     const int* n_dev);

#ifdef __cplusplus
}
#endif
#endif /* D3FEAT_B200_H_ */
