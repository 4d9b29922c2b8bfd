// Final gather of the labels over NVLink peer memory (the only exchange of the path).
//
// Every rank owns one symmetric buffer (mapped into all peers over NVLink / NVSwitch) with one slot per
// batch parity.  flc_scatter_labels_peers is the last kernel of a rank's step: it puts the labels back into
// input order (out[order[i]] = label[i], the scatter the single-GPU path does anyway) straight into the
// rank's slot -- local 4-byte stores; a first version stored into all peers' buffers from this kernel, and
// the random 4-byte NVLink writes cost 0.8 ms at 8 GPUs.  A slot starts with a four-word header
// [n, n_clusters, 0, 0].  After a barrier on the peers' signal pads (issued by the caller on a side stream),
// flc_relabel_gathered PULLS every peer's slot with coalesced 16-byte loads over NVLink and turns them into
// globally unique labels with the running offset of /root/reference/falcon/falcon.py:189-193 (rank r's
// labels + the cluster counts of ranks < r), computed from the headers on the device.  The transfer
// therefore rides on the side stream, off the next batch's critical path, and needs no NCCL kernel.
//
// HBM/NVLink-bound: 4 * n bytes scattered locally; 4 * max_len * world bytes pulled (world - 1 of them
// over NVLink), 4 * max_len * world written.
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
      peers.p[r][slot_offset + 2] = 0;
      peers.p[r][slot_offset + 3] = 0;
    }
  }
  if (i >= n) return;
  const int32_t v = labels[i];
  const int64_t dst = slot_offset + 4 + (order ? static_cast<int64_t>(order[i]) : i);
#pragma unroll 4
  for (int r = 0; r < world; ++r) peers.p[r][dst] = v;
}

struct SlotPtrs {
  const int32_t* p[kMaxPeers];
};

// blockIdx.y = source rank; four labels per thread (16-byte loads; slots are 16-byte aligned, max_len % 4 == 0)
__global__ void relabel_gathered_kernel(SlotPtrs slots, int world, int64_t max_len, int32_t* __restrict__ out,
                                        int64_t* __restrict__ lens) {
  const int64_t i4 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const int r = blockIdx.y;
  int32_t off = 0;
  for (int q = 0; q < r; ++q) off += slots.p[q][1];
  const int32_t len = slots.p[r][0];
  if (i4 == 0) lens[r] = len;
  if (i4 >= max_len) return;
  const int4 v = *reinterpret_cast<const int4*>(slots.p[r] + 4 + i4);
  int4 o;
  o.x = (i4 + 0 < len && v.x >= 0) ? v.x + off : -1;
  o.y = (i4 + 1 < len && v.y >= 0) ? v.y + off : -1;
  o.z = (i4 + 2 < len && v.z >= 0) ? v.z + off : -1;
  o.w = (i4 + 3 < len && v.w >= 0) ? v.w + off : -1;
  *reinterpret_cast<int4*>(out + r * max_len + i4) = o;
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

int flc_relabel_gathered(const void* const* slots, int world, int64_t max_len, int32_t* out, int64_t* lens,
                         flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(world >= 1 && world <= kMaxPeers && max_len >= 0 && (max_len % 4) == 0,
              "bad sizes (at most 16 peers, max_len a multiple of 4)");
  FLC_REQUIRE(slots && out && lens, "null pointer");
  SlotPtrs sp;
  for (int r = 0; r < kMaxPeers; ++r) sp.p[r] = r < world ? static_cast<const int32_t*>(slots[r]) : nullptr;
  for (int r = 0; r < world; ++r)
    FLC_REQUIRE(sp.p[r] != nullptr && (reinterpret_cast<uintptr_t>(sp.p[r]) & 15) == 0, "slots must be 16-byte aligned");
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  cudaStream_t stream = as_stream(stream_);
  const dim3 grid(static_cast<unsigned>((std::max<int64_t>(max_len / 4, 1) + 255) / 256), static_cast<unsigned>(world));
  timed("relabel_gathered", stream, [&] { relabel_gathered_kernel<<<grid, 256, 0, stream>>>(sp, world, max_len, out, lens); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
