// Stable LSD radix sort (8-bit digits) and exclusive scan. Hand-written: no CUB/Thrust.
//
// Per pass: count (per-CTA digit histogram) -> scan (digit-major, CTA-minor) -> scatter (stable local
// ranks from warp match + per-warp counters). The number of passes is fixed by the host-known key width,
// so a sort enqueues a fixed launch sequence with no host round trip.
#include "sort.cuh"

namespace d3f {

__global__ void __launch_bounds__(kSortThreads)
radix_count_kernel(const uint64_t* __restrict__ keys, int Ncap, const int* __restrict__ n_dev, int shift, int nblocks,
                   int* __restrict__ block_hist) {
  const int N = dyn_rows(Ncap, n_dev);
  __shared__ int hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int i = base + r * kSortThreads + threadIdx.x;
    if (i < N) atomicAdd(&hist[(unsigned)(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  block_hist[threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Single-CTA exclusive scan over `n` ints (in place).
__global__ void __launch_bounds__(1024) scan_single_cta_kernel(int* __restrict__ data, int n) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024 * 4) {
    int i0 = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0;
    int tsum = v[0] + v[1] + v[2] + v[3];
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_sums[lane] = winc - w;  // exclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_sums[warp] + inc - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) data[i0 + k] = excl;
      excl += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_sums[31] + inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int Ncap,
                     const int* __restrict__ n_dev, int shift, int nblocks, const int* __restrict__ block_base) {
  const int N = dyn_rows(Ncap, n_dev);
  constexpr int kWarps = kSortThreads / 32;
  __shared__ int warp_cnt[kWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kWarps * 256; d += kSortThreads) (&warp_cnt[0][0])[d] = 0;
  __syncthreads();

  // warp `w` owns the contiguous elements [w*256, (w+1)*256) of the tile; round r = 32 of them in order
  const int base = blockIdx.x * kSortTile + warp * (kSortItems * 32);
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  int rank[kSortItems];
  unsigned dig[kSortItems];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int i = base + r * 32 + lane;
    bool valid = i < N;
    key[r] = valid ? keys_in[i] : 0ull;
    val[r] = valid ? vals_in[i] : 0u;
    dig[r] = valid ? ((unsigned)(key[r] >> shift) & 255u) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
    int leader = __ffs(peers) - 1;
    int old = 0;
    if (valid && lane == leader) {
      old = warp_cnt[warp][dig[r]];
      warp_cnt[warp][dig[r]] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps for each digit, plus the CTA's global base for that digit
  {
    int d = threadIdx.x;  // kSortThreads == 256 digits
    int run = block_base[d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      int c = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    int i = base + r * 32 + lane;
    if (i < N) {
      int pos = warp_cnt[warp][dig[r]] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

int radix_sort_pairs(const SortBuffers& buf, int N, int nbits, cudaStream_t stream, const int* n_dev) {
  if (N <= 0) return 0;
  const int passes = sort_num_passes(nbits);
  const int nblocks = sort_num_blocks(N);
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    int shift = 8 * p;
    radix_count_kernel<<<nblocks, kSortThreads, 0, stream>>>(buf.keys[cur], N, n_dev, shift, nblocks, buf.block_hist);
    D3F_LAUNCH_CHECK("radix_count_kernel");
    scan_single_cta_kernel<<<1, 1024, 0, stream>>>(buf.block_hist, 256 * nblocks);
    D3F_LAUNCH_CHECK("scan_single_cta_kernel");
    radix_scatter_kernel<<<nblocks, kSortThreads, 0, stream>>>(buf.keys[cur], buf.vals[cur], buf.keys[cur ^ 1],
                                                               buf.vals[cur ^ 1], N, n_dev, shift, nblocks,
                                                               buf.block_hist);
    D3F_LAUNCH_CHECK("radix_scatter_kernel");
    cur ^= 1;
  }
  return cur;
}

// ---- exclusive scan over N ints: per-CTA reduce -> single-CTA scan of CTA sums -> per-CTA scan -------
__global__ void __launch_bounds__(256) scan_reduce_kernel(const int* __restrict__ in, int Ncap,
                                                          const int* __restrict__ n_dev,
                                                          int* __restrict__ block_sums) {
  const int N = dyn_rows(Ncap, n_dev);
  __shared__ int ws[8];
  int base = blockIdx.x * 2048;
  int s = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int i = base + r * 256 + threadIdx.x;
    if (i < N) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, int Ncap,
                                                         const int* __restrict__ n_dev,
                                                         const int* __restrict__ block_offsets, int nblocks,
                                                         int* __restrict__ total) {
  const int N = dyn_rows(Ncap, n_dev);
  __shared__ int ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t owns 8 consecutive elements
  int i0 = blockIdx.x * 2048 + threadIdx.x * 8;
  int v[8];
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (i0 + k < N) ? in[i0 + k] : 0;
    tsum += v[k];
  }
  int inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) wbase += ws[w];
  int excl = block_offsets[blockIdx.x] + wbase + inc - tsum;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (i0 + k < N) out[i0 + k] = excl;
    excl += v[k];
  }
  if (total != nullptr && blockIdx.x == nblocks - 1 && threadIdx.x == 255) *total = excl;
}

int exclusive_scan_i32(const int* in, int* out, int N, int* total, int* scratch, cudaStream_t stream,
                       const int* n_dev) {
  if (N <= 0) {
    if (total) D3F_CUDA(cudaMemsetAsync(total, 0, sizeof(int), stream));
    return 0;
  }
  int nblocks = scan_num_blocks(N);
  scan_reduce_kernel<<<nblocks, 256, 0, stream>>>(in, N, n_dev, scratch);
  D3F_LAUNCH_CHECK("scan_reduce_kernel");
  scan_single_cta_kernel<<<1, 1024, 0, stream>>>(scratch, nblocks);
  D3F_LAUNCH_CHECK("scan_single_cta_kernel");
  scan_apply_kernel<<<nblocks, 256, 0, stream>>>(in, out, N, n_dev, scratch, nblocks, total);
  D3F_LAUNCH_CHECK("scan_apply_kernel");
  return 0;
}

}  // namespace d3f
