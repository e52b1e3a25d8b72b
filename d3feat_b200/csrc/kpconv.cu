// KPConv forward (kernels/convolution_ops.py:161-255 rigid, :379-499 deformable).
//
// Stage 1 (this file): one warp per query point gathers the neighbour feature rows (coalesced row reads,
// 128..512 B per warp request), evaluates the kernel-point correlation weights w[h,k] once per neighbour
// (held in shared memory, read back as broadcast LDS.128) and accumulates
//        wf[n,k,:] = sum_h w[n,h,k] * feat[idx[n,h],:]            (:240 / :486)
// in registers. Stage 2 is the dense contraction  out[n,:] = (sum_k wf[n,k,:] @ W[k]) / nn[n]  ==
// [Nq, K*Cin] @ [K*Cin, Cout] with the block epilogue fused (gemm.cu). The [N,H,K,3], [N,H,K], [N,H,Cin]
// intermediates of the TF graph are never materialised; wf is produced in query chunks that stay in L2.
#include "ops.cuh"

namespace d3f {


constexpr int kS1Warps = 4;      // queries per CTA
constexpr int kWStride = 20;     // floats per neighbour in the weight tile (16 B aligned, 4-way write conflicts)
constexpr int kKMax = 16;        // kernel points are padded to 16 in shared memory

struct Stage1Params {
  const float* q;
  const float* s;
  const int* idx;
  const float* feat;
  const unsigned char* flag;  // per support: feature-row sum > 0 (normalisation, :249-253); null if unused
  const float* Kp;            // [K,3]
  const float* offsets;       // [Nq,K,3] or null (deformable)
  const float* modulations;   // [Nq,K] or null
  int Nq, Ns, H, Cin;
  int n0, n1;                 // query chunk [n0, n1)
  float extent;               // KP_extent of this layer
  float inv_scale;            // 1/(2 extent) rigid (:215), 1/extent deformable (:461)
  float gauss_inv;            // 1/(2 sigma^2 + 1e-9), sigma = 0.3 extent (:218-222)
  int influence, closest;
  float shadow;               // coordinate of the shadow support point (1e6 / 1000)
  float* wf;                  // [n1-n0, K*Cin]
  float* inv_nn;              // [n1-n0] or null
};

__global__ void __launch_bounds__(256) rowsum_flag_kernel(const float* __restrict__ feat, int Ns, int Cin,
                                                          unsigned char* __restrict__ flag) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Ns) return;
  float s = 0.f;
  for (int c = lane; c < Cin; c += 32) s += feat[(size_t)warp * Cin + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) flag[warp] = s > 0.f ? 1 : 0;
}

template <int K, int VEC, bool DEFORM>
__global__ void __launch_bounds__(kS1Warps * 32) kpconv_stage1_kernel(Stage1Params p) {
  static_assert(K <= kKMax, "K too large");
  __shared__ __align__(16) float wts[kS1Warps][32 * kWStride];
  __shared__ float kp_s[kS1Warps][kKMax * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n0 + blockIdx.x * kS1Warps + warp;
  if (n >= p.n1) return;  // warp-uniform

  // kernel points of this query (rigid: shared by all queries; deformable: Kp + offsets[n])
  for (int t = lane; t < K * 3; t += 32) {
    float v = p.Kp[t];
    if (DEFORM) v += p.offsets[(size_t)n * K * 3 + t];
    kp_s[warp][t] = v;
  }
  for (int t = K * 3 + lane; t < kKMax * 3; t += 32) kp_s[warp][t] = 0.f;
  const float qx = p.q[3 * (size_t)n], qy = p.q[3 * (size_t)n + 1], qz = p.q[3 * (size_t)n + 2];
  const int* row = p.idx + (size_t)n * p.H;
  const float ext2 = p.extent * p.extent;
  __syncwarp();

  int nn_count = 0;
  const int c_step = 32 * VEC;
  for (int c0 = 0; c0 < p.Cin; c0 += c_step) {
    float acc[K][VEC];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
    const int c = c0 + lane * VEC;
    const bool c_ok = c < p.Cin;

    for (int h0 = 0; h0 < p.H; h0 += 32) {
      // ---- phase A: lane <-> neighbour h0+lane: correlation weights to the K kernel points ------------
      const int h = h0 + lane;
      int id = (h < p.H) ? row[h] : p.Ns;
      if (id < 0 || id > p.Ns) id = p.Ns;  // -1 padding of the non-batch op behaves like the shadow
      const bool real = id < p.Ns;
      float rx, ry, rz;
      if (real) {
        rx = p.s[3 * (size_t)id] - qx; ry = p.s[3 * (size_t)id + 1] - qy; rz = p.s[3 * (size_t)id + 2] - qz;
      } else {
        rx = p.shadow - qx; ry = p.shadow - qy; rz = p.shadow - qz;
      }
      float w[kKMax];
      float dmin = 3.0e38f;
      int kmin = 0;
      bool in_range = false;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float dx = rx - kp_s[warp][3 * k], dy = ry - kp_s[warp][3 * k + 1], dz = rz - kp_s[warp][3 * k + 2];
        float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < dmin) { dmin = d2; kmin = k; }
        in_range = in_range || (d2 < ext2);
        float wk;
        if (p.influence == D3F_INFLUENCE_LINEAR) wk = fmaxf(1.f - sqrtf(d2 + 1e-10f) * p.inv_scale, 0.f);
        else if (p.influence == D3F_INFLUENCE_GAUSSIAN) wk = expf(-d2 * p.gauss_inv);
        else wk = DEFORM ? (d2 < ext2 ? 1.f : 0.f) : 1.f;
        w[k] = wk;
      }
      if (p.closest) {
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = (k == kmin) ? w[k] : 0.f;
      }
      const bool keep = real && (!DEFORM || in_range);
#pragma unroll
      for (int k = 0; k < kKMax; ++k) wts[warp][lane * kWStride + k] = (keep && k < K) ? w[k] : 0.f;
      if (c0 == 0 && p.flag != nullptr) {
        bool cnt = real && p.flag[id] != 0;
        nn_count += __popc(__ballot_sync(0xffffffffu, cnt));
      }
      const unsigned keep_mask = __ballot_sync(0xffffffffu, keep);
      __syncwarp();

      // ---- phase B: accumulate the kept neighbours of this chunk -------------------------------------
      unsigned m = keep_mask;
      while (m) {
        // up to 4 neighbours per step: issue the row loads first (memory-level parallelism)
        int hh[4];
        float f[4][VEC];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (m) {
            hh[u] = __ffs(m) - 1;
            m &= m - 1;
            ++cnt;
          } else {
            hh[u] = -1;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (hh[u] >= 0) {
            int idh = __shfl_sync(0xffffffffu, id, hh[u]);
            const float* fp = p.feat + (size_t)idh * p.Cin + c;
            if (VEC == 4) {
              float4 t = c_ok ? *reinterpret_cast<const float4*>(fp) : make_float4(0.f, 0.f, 0.f, 0.f);
              f[u][0] = t.x; f[u][1 % VEC] = t.y; f[u][2 % VEC] = t.z; f[u][3 % VEC] = t.w;
            } else if (VEC == 2) {
              float2 t = c_ok ? *reinterpret_cast<const float2*>(fp) : make_float2(0.f, 0.f);
              f[u][0] = t.x; f[u][1 % VEC] = t.y;
            } else {
              f[u][0] = c_ok ? *fp : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (hh[u] >= 0) {
            const float4* wp = reinterpret_cast<const float4*>(&wts[warp][hh[u] * kWStride]);
            float wv[kKMax];
#pragma unroll
            for (int kq = 0; kq < (K + 3) / 4; ++kq) {
              float4 t = wp[kq];
              wv[4 * kq] = t.x; wv[4 * kq + 1] = t.y; wv[4 * kq + 2] = t.z; wv[4 * kq + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[k][v] = fmaf(wv[k], f[u][v], acc[k][v]);
          }
        }
        (void)cnt;
      }
      __syncwarp();
    }

    // ---- write wf[n, k, c0 + lane*VEC ..] (optionally modulated, :489-490) ----------------------------
    if (c_ok) {
      float* dst = p.wf + (size_t)(n - p.n0) * K * p.Cin + c;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float mod = (DEFORM && p.modulations) ? p.modulations[(size_t)n * K + k] : 1.f;
        if (VEC == 4) {
          *reinterpret_cast<float4*>(dst + (size_t)k * p.Cin) =
              make_float4(acc[k][0] * mod, acc[k][1 % VEC] * mod, acc[k][2 % VEC] * mod, acc[k][3 % VEC] * mod);
        } else if (VEC == 2) {
          *reinterpret_cast<float2*>(dst + (size_t)k * p.Cin) = make_float2(acc[k][0] * mod, acc[k][1 % VEC] * mod);
        } else {
          dst[(size_t)k * p.Cin] = acc[k][0] * mod;
        }
      }
    }
  }
  if (p.inv_nn != nullptr && lane == 0) p.inv_nn[n - p.n0] = 1.f / (float)max(nn_count, 1);
}

template <int K, bool DEFORM>
static int launch_stage1(const Stage1Params& p, cudaStream_t stream) {
  int nq = p.n1 - p.n0;
  int blocks = ceil_div(nq, kS1Warps);
  bool al16 = (reinterpret_cast<uintptr_t>(p.feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.wf) & 15) == 0;
  if (p.Cin % 128 == 0 && al16) kpconv_stage1_kernel<K, 4, DEFORM><<<blocks, kS1Warps * 32, 0, stream>>>(p);
  else if (p.Cin % 64 == 0 && al16) kpconv_stage1_kernel<K, 2, DEFORM><<<blocks, kS1Warps * 32, 0, stream>>>(p);
  else kpconv_stage1_kernel<K, 1, DEFORM><<<blocks, kS1Warps * 32, 0, stream>>>(p);
  D3F_LAUNCH_CHECK("kpconv_stage1_kernel");
  return D3F_OK;
}

// queries per chunk: keep the wf chunk (K*Cin floats per query) around 48 MB so it is produced and
// consumed out of the 126 MB L2 instead of HBM
static int chunk_queries(int K, int Cin) {
  long long per = (long long)K * Cin * 4;
  long long n = (48ll << 20) / per;
  if (n < 1024) n = 1024;
  if (n > (1 << 20)) n = 1 << 20;
  return (int)(n / 128 * 128);
}

size_t kpconv_workspace_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout) {
  (void)H; (void)Cout;
  int chunk = chunk_queries(K, Cin);
  if (chunk > Nq) chunk = Nq > 0 ? Nq : 1;
  size_t b = 0;
  b += align_up((size_t)chunk * K * Cin * sizeof(float), 256);
  b += align_up((size_t)chunk * sizeof(float), 256);
  b += align_up((size_t)(Ns + 1), 256);
  return b + 1024;
}

int kpconv_forward_impl(bool deform, const float* q, const float* s, const int* idx, const float* feat,
                        const float* Kp, const float* offsets, const float* modulations, const float* W, int Nq,
                        int Ns, int H, int K, int Cin, int Cout, float extent, int influence, int mode, int normalize,
                        const float* bn_scale, const float* bn_shift, const float* bias, float leaky_alpha,
                        float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  D3F_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 0 && Cin >= 1 && Cout >= 1, D3F_ERR_INVALID,
              "kpconv: bad shape Nq=%d Ns=%d H=%d Cin=%d Cout=%d", Nq, Ns, H, Cin, Cout);
  D3F_REQUIRE(K == 15, D3F_ERR_INVALID, "kpconv: num_kernel_points=%d not instantiated (built for K=15)", K);
  D3F_REQUIRE(influence >= 0 && influence <= 2, D3F_ERR_INVALID,
              "Unknown influence function type (config.KP_influence)");
  D3F_REQUIRE(mode == D3F_MODE_SUM || mode == D3F_MODE_CLOSEST, D3F_ERR_INVALID,
              "Unknown convolution mode. Should be 'closest' or 'sum'");
  D3F_REQUIRE(extent > 0.f, D3F_ERR_INVALID, "kpconv: KP_extent=%g", (double)extent);
  D3F_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), D3F_ERR_INVALID, "kpconv: bn_scale/bn_shift mismatch");
  D3F_REQUIRE(!deform || offsets != nullptr, D3F_ERR_INVALID, "kpconv_deform: offsets missing");
  D3F_REQUIRE(workspace_bytes >= kpconv_workspace_bytes(Nq, Ns, H, K, Cin, Cout), D3F_ERR_WORKSPACE,
              "kpconv: workspace too small");
  if (Nq == 0) return D3F_OK;
  int chunk = chunk_queries(K, Cin);
  if (chunk > Nq) chunk = Nq;
  Carver cv(workspace, workspace_bytes);
  float* wf = cv.take<float>((size_t)chunk * K * Cin);
  float* inv_nn = cv.take<float>(chunk);
  unsigned char* flag = cv.take<unsigned char>(Ns + 1);

  const bool norm = normalize != 0 && !deform;
  if (norm && Ns > 0) {
    rowsum_flag_kernel<<<ceil_div(Ns * 32, 256), 256, 0, stream>>>(feat, Ns, Cin, flag);
    D3F_LAUNCH_CHECK("rowsum_flag_kernel");
  }
  Stage1Params p;
  p.q = q; p.s = s; p.idx = idx; p.feat = feat; p.flag = norm ? flag : nullptr;
  p.Kp = Kp; p.offsets = offsets; p.modulations = modulations;
  p.Nq = Nq; p.Ns = Ns; p.H = H; p.Cin = Cin;
  p.extent = extent;
  p.inv_scale = deform ? 1.f / extent : 1.f / (2.f * extent);
  float sigma = extent * 0.3f;
  p.gauss_inv = 1.f / (2.f * sigma * sigma + 1e-9f);
  p.influence = influence;
  p.closest = mode == D3F_MODE_CLOSEST;
  p.shadow = deform ? 1000.f : 1e6f;
  p.wf = wf;
  p.inv_nn = norm ? inv_nn : nullptr;
  for (int n0 = 0; n0 < Nq; n0 += chunk) {
    p.n0 = n0;
    p.n1 = min(Nq, n0 + chunk);
    int rc = deform ? launch_stage1<15, true>(p, stream) : launch_stage1<15, false>(p, stream);
    if (rc) return rc;
    Epilogue ep;
    ep.rowscale = norm ? inv_nn : nullptr;
    ep.bn_scale = bn_scale; ep.bn_shift = bn_shift; ep.bias = bias; ep.residual = nullptr;
    ep.leaky_alpha = leaky_alpha;
    rc = gemm_f32(wf, W, out + (size_t)n0 * Cout, p.n1 - n0, Cout, K * Cin, ep, stream);
    if (rc) return rc;
  }
  return D3F_OK;
}

}  // namespace d3f
