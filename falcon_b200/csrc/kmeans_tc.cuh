// k-means assignment of large ("tiled") buckets on tcgen05 tensor cores (kmeans_tc.cu).
#pragma once
#include "common.cuh"

namespace flc {

// units[u] = (first row, bucket end row, first centroid row, end centroid row) of one 128-row
// query tile; *n_units on the device.  best[i] = arg-max list of row i by bf16 scores; rows
// whose two best scores are closer than `margin` are appended to unsure_list (length
// *n_unsure, on the device): the caller re-scores those exactly.
int launch_kmeans_tc(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, const uint16_t* c_bf16, int64_t ld_c,
                     int64_t total_centroids, uint32_t low_dim, const int4* units, const int32_t* n_units,
                     float margin, int32_t* best, int32_t* unsure_list, int32_t* n_unsure, cudaStream_t stream);

}  // namespace flc
