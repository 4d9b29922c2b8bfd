// k-means assignment for buckets too large for the shared-memory trainer, on tcgen05
// tensor cores: S = X_tile * C_b^T (128 rows x up to 256 lists per MMA tile, K =
// low_dim) from the bf16 rows the scan already uses and a bf16 copy of the bucket's
// centroids.  The epilogue keeps the two best scores of every row; a row whose two
// best lists are closer than the margin (2^-7 covers twice the bf16 rounding error of
// unit vectors) is flagged and re-scored exactly by the caller, so the assignment is
// the float32 arg-max of the reference arithmetic -- the tensor cores only prove, for
// most rows, which list that is.
//
// Same structure as scan_tc.cu: persistent grid (1 CTA/SM), warp 0 = TMA producer
// (128x64 boxes, SWIZZLE_128B, 4-stage mbarrier ring), warp 1 = MMA issuer
// (tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators double-buffered in TMEM),
// warps 2-5 = epilogue (tcgen05.ld 32x32b.x32 -> running top-2 per row).
// HBM-bound: every bf16 row is read once per iteration (2 * ld bytes per row).
#include "kmeans_tc.cuh"
#include "tc_common.cuh"

namespace flc {

struct KTile {
  int q0, q_end, c0, c_end;
  bool first, last;  // first / last list tile of the unit
};

struct UnitWalker {
  const int4* units;
  int64_t total, stride, u;
  int4 cur, nxt;
  int ct, tiles_c;

  __device__ static int count_tiles(const int4& d) { return max(1, (d.w - d.z + kTileN - 1) / kTileN); }
  __device__ void init(const int4* ud, int64_t n_units, int64_t first, int64_t step) {
    units = ud; total = n_units; stride = step; u = first; ct = 0;
    cur = nxt = make_int4(0, 0, 0, 0);
    if (u < total) cur = __ldg(units + u);
    if (u + stride < total) nxt = __ldg(units + u + stride);
    tiles_c = count_tiles(cur);
  }
  __device__ bool valid() const { return u < total; }
  __device__ KTile get() const {
    KTile t;
    t.q0 = cur.x; t.q_end = cur.y;
    t.c0 = cur.z + ct * kTileN; t.c_end = cur.w;
    t.first = ct == 0; t.last = ct == tiles_c - 1;
    return t;
  }
  __device__ void next() {
    if (++ct >= tiles_c) {
      ct = 0;
      u += stride;
      cur = nxt;
      tiles_c = count_tiles(cur);
      if (u + stride < total) nxt = __ldg(units + u + stride);
    }
  }
};

__global__ void __launch_bounds__(kScanThreads, 1)
kmeans_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_c,
                 uint32_t low_dim, const int4* __restrict__ units, const int32_t* __restrict__ n_units_ptr,
                 float margin, int32_t* __restrict__ best_out, int32_t* __restrict__ unsure_list,
                 int32_t* __restrict__ n_unsure) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles_base = (raw + 1023u) & ~1023u;
  const uint32_t bar_base = tiles_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int num_kb = static_cast<int>((low_dim + kBoxK - 1) / kBoxK);
  const int64_t n_units = *n_units_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      UnitWalker w;
      int stage = 0;
      uint32_t phase = 0;
      for (w.init(units, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
        const KTile t = w.get();
        const bool two = (t.c_end - t.c0) > kBoxRows;
        const uint32_t bytes = kABytes + (two ? 2 : 1) * kBoxBytes;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), bytes);
          const uint32_t sa = tiles_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          tma_load_2d(sa, &tmap_x, kb * kBoxK, t.q0, full_bar(stage));
          tma_load_2d(sb, &tmap_c, kb * kBoxK, t.c0, full_bar(stage));
          if (two) tma_load_2d(sb + kBoxBytes, &tmap_c, kb * kBoxK, t.c0 + kBoxRows, full_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    UnitWalker w;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (w.init(units, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
      const KTile t = w.get();
      const int nc = min(kTileN, t.c_end - t.c0);
      const uint32_t n_mma = static_cast<uint32_t>(max(16, (nc + 15) & ~15));
      const uint32_t idesc = make_idesc(kTileM, n_mma);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kTileN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = tiles_base + stage * kStageBytes;
          const uint32_t sb = sa + kABytes;
          const int rem = static_cast<int>(low_dim) - kb * kBoxK;
          const int ksteps = rem >= kBoxK ? kBoxK / 16 : (rem + 15) / 16;
          for (int k = 0; k < ksteps; ++k)
            tc_mma_bf16(d_tmem, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (kb == num_kb - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue (warps 2..5): running top-2 of every row =====================
    const int quarter = warp & 3;  // TMEM lanes [32 * quarter, 32 * quarter + 32)
    UnitWalker w;
    int acc = 0;
    uint32_t acc_phase = 0;
    float best = -INFINITY, second = -INFINITY;
    int best_id = 0;
    for (w.init(units, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
      const KTile t = w.get();
      const int nc = min(kTileN, t.c_end - t.c0);
      const int q = t.q0 + quarter * 32 + lane;
      if (t.first) { best = -INFINITY; second = -INFINITY; best_id = 0; }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int chunks = (nc + 31) >> 5;
      const int id0 = t.c0 - w.cur.z;  // list id of the tile's first column
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kTileN + ch * 32);
        tc_ld_32x32(taddr, v);
        const int cols = min(32, nc - ch * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = __uint_as_float(v[j]);
          if (j < cols) {
            if (s > best) { second = best; best = s; best_id = id0 + ch * 32 + j; }
            else if (s > second) { second = s; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      if (t.last) {
        const bool have = q < t.q_end;
        if (have) best_out[q] = best_id;
        // rows the bf16 scores cannot decide (a single list: second = -inf, decided) go on the
        // re-score list: one atomic per warp
        const bool unsure = have && !(best - second > margin);
        const uint32_t ub = __ballot_sync(0xffffffffu, unsure);
        if (ub != 0u) {
          int base = 0;
          if (lane == 0) base = atomicAdd(n_unsure, __popc(ub));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (unsure) unsure_list[base + __popc(ub & ((1u << lane) - 1u))] = q;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                 : "memory");
  }
}

int launch_kmeans_tc(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, const uint16_t* c_bf16, int64_t ld_c,
                     int64_t total_centroids, uint32_t low_dim, const int4* units, const int32_t* n_units,
                     float margin, int32_t* best, int32_t* unsure_list, int32_t* n_unsure, cudaStream_t stream) {
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(c_bf16) & 15) == 0,
              "bf16 matrices must be 16-byte aligned");
  FLC_REQUIRE((ld_bf16 % 8) == 0 && (ld_c % 8) == 0, "bf16 row pitches must be multiples of 8");
  CUtensorMap tmap_x, tmap_c;
  FLC_TRY(make_bf16_tmap(&tmap_x, x_bf16, static_cast<uint64_t>(n), low_dim, ld_bf16));
  FLC_TRY(make_bf16_tmap(&tmap_c, c_bf16, static_cast<uint64_t>(total_centroids), low_dim, ld_c));
  static bool attr_set = false;
  if (!attr_set) {
    FLC_CUDA(cudaFuncSetAttribute(kmeans_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  timed("kmeans_tc", stream, [&] { kmeans_tc_kernel<<<kNumSMs, kScanThreads, kSmemBytes, stream>>>(
      tmap_x, tmap_c, low_dim, units, n_units, margin, best, unsure_list, n_unsure); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // namespace flc
