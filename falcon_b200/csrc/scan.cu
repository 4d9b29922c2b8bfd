// Stage 3a host side: tile schedule + dispatch of the inverted-list scan, and the
// SIMT verification kernel (impl = 1).  The production kernel (impl = 0) is the
// tcgen05/TMA kernel in scan_tc.cu.
#include <cub/cub.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace flc {

// ---------------------------------------------------------------- tile schedule
// Bucket b with n_b rows owns ceil(n_b / kTileM) query tiles ("units"); a unit is
// scanned against the bucket's ceil(n_b / kTileN) candidate tiles.  tile_off =
// exclusive scan of the unit counts over buckets.
__global__ void tile_count_kernel(const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                                  int64_t* __restrict__ tiles) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;
  if (b == n_buckets) {
    tiles[b] = 0;
    return;
  }
  const int64_t nb = bucket_ptr[b + 1] - bucket_ptr[b];
  tiles[b] = (nb + kTileM - 1) / kTileM;
}

// ---------------------------------------------------------------- SIMT verification kernel
// One warp per query; bf16 inputs, float32 accumulation.  Same contract as the
// tensor-core kernel: every within-bucket pair with ip >= threshold is emitted.
__global__ void __launch_bounds__(128)
scan_simt_kernel(const uint16_t* __restrict__ x, int64_t ld, int64_t n, uint32_t low_dim,
                 const int64_t* __restrict__ bucket_ptr, int64_t n_buckets, float threshold,
                 uint64_t* __restrict__ pairs, uint64_t capacity,
                 unsigned long long* __restrict__ pair_count) {
  extern __shared__ float smem_q[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* xq = smem_q + static_cast<size_t>(warp) * low_dim;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (q >= n) return;
  // bucket of q: last b with bucket_ptr[b] <= q
  int64_t lo = 0, hi = n_buckets;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (bucket_ptr[mid] <= q) lo = mid; else hi = mid;
  }
  const int64_t c0 = bucket_ptr[lo], c1 = bucket_ptr[lo + 1];
  for (uint32_t i = lane; i < low_dim; i += 32)
    xq[i] = __uint_as_float(static_cast<uint32_t>(x[q * ld + i]) << 16);
  __syncwarp();
  for (int64_t c = c0; c < c1; ++c) {
    const uint16_t* xc = x + c * ld;
    float acc = 0.f;
    for (uint32_t i = lane; i < low_dim; i += 32)
      acc = fmaf(xq[i], __uint_as_float(static_cast<uint32_t>(__ldg(xc + i)) << 16), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc >= threshold) {
      const unsigned long long slot = atomicAdd(pair_count, 1ull);
      if (slot < capacity) pairs[slot] = (static_cast<uint64_t>(q) << 32) | static_cast<uint64_t>(c);
    }
  }
}

// Unit descriptors (first query row, first candidate row, bucket end, candidate tiles): one warp per
// bucket writes those of the bucket's units.  S = X X^T is symmetric, so a query tile is only scanned
// against the candidate rows from its own first row on (candidate tiles start there: TMA boxes may
// start at any row); the kernel emits both (q, c) and (c, q) for every hit above the diagonal.
// About half the tiles of a large bucket disappear, and the second query tile of a small bucket
// meets only the bucket's tail.
__global__ void unit_desc_kernel(const int64_t* __restrict__ bucket_ptr, const int64_t* __restrict__ tile_off,
                                 int64_t n_buckets, int4* __restrict__ unit_desc) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (b >= n_buckets) return;
  const int64_t s = bucket_ptr[b], e = bucket_ptr[b + 1], u0 = tile_off[b];
  // Position p of the bucket's descriptor list holds query tile p / 2 from the front (p even) or from the
  // back (p odd): a query tile's cost falls linearly with its index (it only meets the candidate rows from
  // its own first row on), so every consecutive PAIR of positions costs the same, and the kernel deals
  // positions to CTAs in pairs.
  const int64_t tiles_q = tile_off[b + 1] - u0;
  for (int64_t p = lane; p < tiles_q; p += 32) {
    const int64_t i = (p & 1) ? tiles_q - 1 - (p >> 1) : (p >> 1);
    const int64_t q_rel = i * kTileM;
    const int64_t c_first = s + q_rel;
    const int tiles_c = static_cast<int>((e - c_first + kTileN - 1) / kTileN);
    unit_desc[u0 + p] = make_int4(static_cast<int>(s + q_rel), static_cast<int>(c_first), static_cast<int>(e), tiles_c);
  }
}

struct ScanLayout {
  int64_t* tiles;        // [n_buckets + 1]
  int64_t* tile_off;     // [n_buckets + 1]
  int4* unit_desc;       // [n / kTileM + n_buckets + 1]
  void* cub_tmp;
  size_t cub_bytes;
};

static void scan_layout(Workspace& ws, int64_t n, int64_t n_buckets, ScanLayout& L) {
  L.tiles = ws.take<int64_t>(n_buckets + 1);
  L.tile_off = ws.take<int64_t>(n_buckets + 1);
  L.unit_desc = ws.take<int4>(static_cast<size_t>(n / kTileM + n_buckets + 1));
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int64_t*)nullptr, (int64_t*)nullptr,
                                static_cast<int>(n_buckets + 1));
  L.cub_bytes = b;
  L.cub_tmp = ws.take<char>(b);
}

}  // namespace flc

extern "C" {

size_t flc_scan_workspace_bytes(int64_t n, int64_t n_buckets) {
  if (n_buckets <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::ScanLayout L;
  flc::scan_layout(ws, n > 0 ? n : 0, n_buckets, L);
  return ws.used + 256;
}

int flc_scan_pairs(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, uint32_t low_dim,
                   const int64_t* bucket_ptr, int64_t n_buckets, const int32_t* list_id,
                   const int32_t* probes, int32_t max_nprobe, const int32_t* nlist, float threshold,
                   int impl, uint64_t* pairs, uint64_t pair_capacity, uint64_t* pair_count,
                   void* workspace, size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  (void)list_id; (void)probes; (void)max_nprobe; (void)nlist;  // tile skipping: later round
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "n out of range");
  FLC_REQUIRE(impl == 0 || impl == 1, "impl must be 0 (tcgen05) or 1 (simt)");
  FLC_REQUIRE(pair_count != nullptr, "null pair_count");
  cudaStream_t stream = as_stream(stream_);
  FLC_CUDA(cudaMemsetAsync(pair_count, 0, sizeof(uint64_t), stream));
  if (n == 0 || n_buckets == 0) return FLC_OK;
  FLC_REQUIRE(ld_bf16 >= low_dim && (ld_bf16 % 8) == 0, "ld_bf16 must be >= low_dim and a multiple of 8");
  if (impl == 1) {
    const size_t smem = static_cast<size_t>(4) * low_dim * sizeof(float);
    if (smem > 48 * 1024)
      FLC_CUDA(cudaFuncSetAttribute(scan_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    timed("scan_simt", stream, [&] { scan_simt_kernel<<<static_cast<unsigned>((n + 3) / 4), 128, smem, stream>>>(
        x_bf16, ld_bf16, n, low_dim, bucket_ptr, n_buckets, threshold, pairs, pair_capacity,
        reinterpret_cast<unsigned long long*>(pair_count)); });
    FLC_LAUNCH_CHECK();
    return FLC_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  ScanLayout L;
  scan_layout(ws, n, n_buckets, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "scan workspace too small: need %zu", ws.used);
  timed("tile_count", stream, [&] { tile_count_kernel<<<static_cast<unsigned>((n_buckets + 1 + 255) / 256), 256, 0, stream>>>(
      bucket_ptr, n_buckets, L.tiles); });
  FLC_LAUNCH_CHECK();
  size_t tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.tiles, L.tile_off,
                                         static_cast<int>(n_buckets + 1), stream));
  count_launch(2);
  timed("unit_desc", stream, [&] { unit_desc_kernel<<<static_cast<unsigned>((n_buckets * 32 + 255) / 256), 256, 0, stream>>>(
      bucket_ptr, L.tile_off, n_buckets, L.unit_desc); });
  FLC_LAUNCH_CHECK();
  return launch_scan_tc(x_bf16, ld_bf16, n, low_dim, bucket_ptr, n_buckets, L.tile_off, L.unit_desc, threshold, pairs,
                        pair_capacity, reinterpret_cast<unsigned long long*>(pair_count), stream);
}

}  // extern "C"
