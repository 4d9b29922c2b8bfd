// Final gather of the labels over NVLink peer memory (the only exchange of the path).
//
// Every rank owns one symmetric buffer of `world` slots per batch parity; slot r of EVERY rank's buffer
// receives rank r's labels.  flc_scatter_labels_peers is the last kernel of a rank's step: it puts the
// labels back into input order (out[order[i]] = label[i], the scatter the single-GPU path does anyway) and
// stores each value straight into slot `rank` of all peers' buffers -- NVLink stores from the producing
// kernel, no separate collective, no staging copy.  A slot starts with a two-word header [n, n_clusters].
// After a barrier on the peers' signal pads (issued by the caller on a side stream),
// flc_relabel_gathered turns the raw slots into globally unique labels with the running offset of
// /root/reference/falcon/falcon.py:189-193 (rank r's labels + the cluster counts of ranks < r), computed
// from the headers on the device.
//
// HBM/NVLink-bound: 4 * n bytes read twice (labels, order), 4 * n * world bytes stored.
#include "common.cuh"

namespace flc {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  int32_t* p[kMaxPeers];
};

__global__ void scatter_labels_peers_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ order,
                                            int64_t n, const int64_t* __restrict__ n_clusters_dev,
                                            int64_t n_clusters, PeerPtrs peers, int world, int64_t slot_offset) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) {
    const int32_t nc = static_cast<int32_t>(n_clusters_dev ? *n_clusters_dev : n_clusters);
    for (int r = 0; r < world; ++r) {
      peers.p[r][slot_offset] = static_cast<int32_t>(n);
      peers.p[r][slot_offset + 1] = nc;
    }
  }
  if (i >= n) return;
  const int32_t v = labels[i];
  const int64_t dst = slot_offset + 2 + (order ? static_cast<int64_t>(order[i]) : i);
#pragma unroll 4
  for (int r = 0; r < world; ++r) peers.p[r][dst] = v;
}

__global__ void relabel_gathered_kernel(const int32_t* __restrict__ slots, int world, int64_t max_len,
                                        int32_t* __restrict__ out, int64_t* __restrict__ lens) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const int64_t slot = max_len + 2;
  int32_t off = 0;
  for (int q = 0; q < r; ++q) off += slots[q * slot + 1];
  const int32_t len = slots[r * slot];
  if (i == 0) lens[r] = len;
  if (i >= max_len) return;
  const int32_t v = i < len ? slots[r * slot + 2 + i] : -1;
  out[r * max_len + i] = v >= 0 ? v + off : -1;
}

}  // namespace flc

extern "C" {

int flc_scatter_labels_peers(const int32_t* labels, const int32_t* order, int64_t n,
                             const int64_t* n_clusters_dev, int64_t n_clusters, void* const* peer_buffers,
                             int world, int64_t slot_offset, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && world >= 1 && world <= kMaxPeers, "bad sizes (at most 16 peers)");
  FLC_REQUIRE(peer_buffers != nullptr && slot_offset >= 0, "null peer buffers");
  PeerPtrs peers;
  for (int r = 0; r < kMaxPeers; ++r) peers.p[r] = r < world ? static_cast<int32_t*>(peer_buffers[r]) : nullptr;
  for (int r = 0; r < world; ++r) FLC_REQUIRE(peers.p[r] != nullptr, "null peer buffer");
  cudaStream_t stream = as_stream(stream_);
  const unsigned blocks = static_cast<unsigned>((std::max<int64_t>(n, 1) + 255) / 256);
  timed("scatter_labels_peers", stream, [&] { scatter_labels_peers_kernel<<<blocks, 256, 0, stream>>>(
      labels, order, n, n_clusters_dev, n_clusters, peers, world, slot_offset); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

int flc_relabel_gathered(const int32_t* slots, int world, int64_t max_len, int32_t* out, int64_t* lens,
                         flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(world >= 1 && world <= kMaxPeers && max_len >= 0, "bad sizes (at most 16 peers)");
  FLC_REQUIRE(slots && out && lens, "null pointer");
  cudaStream_t stream = as_stream(stream_);
  const dim3 grid(static_cast<unsigned>((std::max<int64_t>(max_len, 1) + 255) / 256), static_cast<unsigned>(world));
  timed("relabel_gathered", stream, [&] { relabel_gathered_kernel<<<grid, 256, 0, stream>>>(slots, world, max_len, out, lens); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
