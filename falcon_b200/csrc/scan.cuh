// Shared declarations of the inverted-list scan (stage 3a).
#pragma once
#include "common.cuh"

namespace flc {

constexpr int kTileM = 128;  // queries per tile   (UMMA M, one TMEM lane per query)
constexpr int kTileN = 256;  // candidates per tile (max UMMA N, one TMEM column each)
constexpr int kBoxK = 64;    // bf16 per TMA box row = 128 bytes = one SWIZZLE_128B span

// tcgen05/TMA kernel launcher (scan_tc.cu).  tile_off is the exclusive scan of the
// per-bucket query-tile ("unit") counts (length n_buckets + 1, last entry = total).
int launch_scan_tc(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, uint32_t low_dim,
                   const int64_t* bucket_ptr, int64_t n_buckets, const int64_t* tile_off,
                   const int4* unit_desc /*[total units]: first query row, bucket start, bucket end, candidate tiles*/,
                   float threshold, uint64_t* pairs, uint64_t pair_capacity,
                   unsigned long long* pair_count, cudaStream_t stream);

}  // namespace flc
