// k-means assignment of large ("tiled") buckets on tcgen05 tensor cores (kmeans_tc.cu).
#pragma once
#include "common.cuh"

namespace flc {

// units[u] = (first row, bucket end row, first centroid row, end centroid row) of one 128-row
// query tile; *n_units on the device.  best[i] = arg-max list of row i by bf16 scores; rows
// whose two best scores are closer than `margin` have the top bit of best[i] set
// (kTcUnsure): the caller re-scores those exactly.
constexpr int32_t kTcUnsure = static_cast<int32_t>(0x80000000u);
int launch_kmeans_tc(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, const uint16_t* c_bf16, int64_t ld_c,
                     int64_t total_centroids, uint32_t low_dim, const int4* units, const int32_t* n_units,
                     float margin, int32_t* best, cudaStream_t stream);

// The same assignment for buckets of up to kSparseMaxLists lists, from the SPARSE rows: builder
// warps expand each 128-row tile into the swizzled bf16 operand layout in shared memory (zero fill +
// scatter of the row's populated columns), so a row costs its ~200 sparse bytes of HBM traffic instead
// of the 2 * ld bytes of its dense bf16 copy.  Needs low_dim <= kSparseMaxDim and ell_width <=
// kSparseMaxWidth (kmeans_tc_sparse_ok); units are walked in contiguous ranges per CTA so that a
// bucket's centroids stay resident in shared memory across its tiles.  *neg_seen (device) is set when
// a scattered value has its sign bit set; with use_neg_seen != 0 and *neg_seen == 0 on entry (from an
// earlier launch over the same rows) the kernel uses the tighter relative margin valid for
// non-negative operands instead of `margin`.
constexpr int kSparseMaxLists = 64;
constexpr int kSparseMaxDim = 448;
constexpr int kSparseMaxWidth = 64;
inline bool kmeans_tc_sparse_ok(uint32_t low_dim, int32_t ell_width) {
  return low_dim <= kSparseMaxDim && ell_width <= kSparseMaxWidth && (ell_width % 8) == 0;
}
int launch_kmeans_tc_sparse(const uint16_t* ell_idx, const float* ell_val, const uint16_t* ell_nnz, int32_t ell_width,
                            const uint16_t* c_bf16, int64_t ld_c, int64_t total_centroids, uint32_t low_dim,
                            const int4* units, const int32_t* n_units, float margin, int32_t* neg_seen,
                            int use_neg_seen, int32_t* best, cudaStream_t stream);

}  // namespace flc
