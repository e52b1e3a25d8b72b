// Device-wide primitives used by the hash-grid ops: stable LSD radix sort of (u64 key, u32 value)
// pairs on a host-known number of key bits, and an exclusive int32 scan.
#pragma once
#include "common.cuh"

namespace d3f {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per CTA

struct SortBuffers {
  uint64_t* keys[2];
  uint32_t* vals[2];
  int* block_hist;  // [256 * nblocks]
};

inline int sort_num_blocks(int N) { return N > 0 ? ceil_div(N, kSortTile) : 1; }
inline int sort_num_passes(int nbits) { return nbits <= 0 ? 0 : (nbits + 7) / 8; }

// N is the launch capacity; with n_dev the kernels read the actual element count from device memory (<= N).
// Sorts by the low `nbits` bits. Input in buffers[0]; returns the index (0/1) of the buffer that holds
// the sorted pairs (== passes & 1, host-known), or a negative error code.
int radix_sort_pairs(const SortBuffers& buf, int N, int nbits, cudaStream_t stream, const int* n_dev = nullptr);

// out[i] = sum_{j<i} in[j]; total[0] = sum of all (optional). scratch: ceil(N/2048)+1 ints.
inline int scan_num_blocks(int N) { return N > 0 ? ceil_div(N, 2048) : 1; }
int exclusive_scan_i32(const int* in, int* out, int N, int* total, int* scratch, cudaStream_t stream,
                       const int* n_dev = nullptr);

}  // namespace d3f
