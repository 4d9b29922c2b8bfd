// Cluster representatives (medoids) from the sparse k-NN matrix -- SURVEY 8a row
// a16 / A.5 (published falcon `get_cluster_representatives`; the snapshot's dense
// descendant is falcon/cluster/cluster.py:512-553).  For every non-noise cluster:
//   * at most two members: the first member (all pairwise distances are equal);
//   * otherwise the member whose sparse row has the smallest mean distance to the
//     cluster members present in that row (the row's own zero-distance entry
//     included), a row being eligible only if more than a quarter of the cluster
//     is present in it; ties and "no eligible row" resolve to the first member.
// The mean is a float32 sum in stored row order divided by the float32 count.
//
// One thread per row (rows hold at most n_neighbors entries); the arg-min per
// cluster is one 64-bit atomicMin on (mean bits << 32 | row).  HBM-bound:
// nnz * 12 (dist + index + label gather) + n * 12 bytes.
#include "common.cuh"

namespace flc {

__global__ void medoid_size_kernel(const int32_t* __restrict__ labels, int64_t n, int64_t n_clusters,
                                   int32_t* __restrict__ size) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t l = labels[i];
  if (l >= 0 && l < n_clusters) atomicAdd(size + l, 1);
}

__global__ void medoid_row_kernel(const float* __restrict__ dist, const int32_t* __restrict__ indices,
                                  const int64_t* __restrict__ indptr, int64_t n, const int32_t* __restrict__ labels,
                                  int64_t n_clusters, const int32_t* __restrict__ size,
                                  unsigned long long* __restrict__ best) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t l = labels[i];
  if (l < 0 || l >= n_clusters) return;
  const int32_t sz = size[l];
  uint32_t key_hi = 0u;  // clusters of one or two members: every row ties, the first member wins
  if (sz > 2) {
    float sum = 0.f;
    int32_t cnt = 0;
    for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) {
      const int32_t c = __ldg(indices + p);
      if (c >= 0 && c < n && __ldg(labels + c) == l) {
        sum += __ldg(dist + p);
        ++cnt;
      }
    }
    // eligible only if more than a quarter of the cluster is present in the row
    const float avg = (4 * static_cast<int64_t>(cnt) > sz) ? sum / static_cast<float>(cnt) : INFINITY;
    key_hi = __float_as_uint(fmaxf(avg, 0.f));  // non-negative floats order like their bit patterns
  }
  atomicMin(best + l, (static_cast<unsigned long long>(key_hi) << 32) | static_cast<unsigned long long>(i));
}

__global__ void medoid_emit_kernel(const unsigned long long* __restrict__ best, int64_t n_clusters,
                                   int32_t* __restrict__ medoids) {
  const int64_t l = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (l >= n_clusters) return;
  const unsigned long long k = best[l];
  medoids[l] = k == ~0ull ? -1 : static_cast<int32_t>(k & 0xffffffffull);
}

}  // namespace flc

extern "C" {

size_t flc_medoids_workspace_bytes(int64_t n_clusters) {
  const size_t c = static_cast<size_t>(n_clusters > 0 ? n_clusters : 1);
  return flc::align256(c * sizeof(int32_t)) + flc::align256(c * sizeof(unsigned long long)) + 256;
}

int flc_medoids(const float* dist, const int32_t* indices, const int64_t* indptr, int64_t n,
                const int32_t* labels, int64_t n_clusters, int32_t* medoids, void* workspace,
                size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31) && n_clusters >= 0, "bad sizes");
  if (n_clusters == 0) return FLC_OK;
  FLC_REQUIRE(dist && indices && indptr && labels && medoids, "null argument");
  cudaStream_t stream = as_stream(stream_);
  Workspace ws(workspace, workspace_bytes);
  int32_t* size = ws.take<int32_t>(n_clusters);
  unsigned long long* best = ws.take<unsigned long long>(n_clusters);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "medoids workspace too small: need %zu", ws.used);
  FLC_CUDA(cudaMemsetAsync(size, 0, sizeof(int32_t) * n_clusters, stream));
  FLC_CUDA(cudaMemsetAsync(best, 0xff, sizeof(unsigned long long) * n_clusters, stream));
  const unsigned rb = static_cast<unsigned>((n + 255) / 256), cb = static_cast<unsigned>((n_clusters + 255) / 256);
  if (n > 0) {
    timed("medoid_size", stream, [&] { medoid_size_kernel<<<rb, 256, 0, stream>>>(labels, n, n_clusters, size); });
    FLC_LAUNCH_CHECK();
    timed("medoid_row", stream, [&] { medoid_row_kernel<<<rb, 256, 0, stream>>>(
        dist, indices, indptr, n, labels, n_clusters, size, best); });
    FLC_LAUNCH_CHECK();
  }
  timed("medoid_emit", stream, [&] { medoid_emit_kernel<<<cb, 256, 0, stream>>>(best, n_clusters, medoids); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
