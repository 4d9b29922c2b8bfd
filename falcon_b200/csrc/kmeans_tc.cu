// k-means assignment for buckets too large for the shared-memory trainer, on tcgen05
// tensor cores: S = X_tile * C_b^T (128 rows x up to 256 lists per MMA tile, K =
// low_dim) from the bf16 rows the scan already uses and a bf16 copy of the bucket's
// centroids.  The epilogue keeps the two best scores of every row; a row whose two
// best lists are closer than the margin (2^-7 covers twice the bf16 rounding error of
// unit vectors) is flagged (top bit of its entry) and re-scored exactly by the caller, so the assignment is
// the float32 arg-max of the reference arithmetic -- the tensor cores only prove, for
// most rows, which list that is.
//
// Same structure as scan_tc.cu: persistent grid (1 CTA/SM), warp 0 = TMA producer
// (128x64 boxes, SWIZZLE_128B, 4-stage mbarrier ring), warp 1 = MMA issuer
// (tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators double-buffered in TMEM),
// warps 2-5 = epilogue (tcgen05.ld 32x32b.x32 -> running top-2 per row).
// HBM-bound: every bf16 row is read once per iteration (2 * ld bytes per row).
#include <cuda_bf16.h>

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
                 float margin, int32_t* __restrict__ best_out) {
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
        // rows the bf16 scores cannot decide (a single list: second = -inf, decided) are flagged for the
        // exact re-score
        if (q < t.q_end) best_out[q] = (best - second > margin) ? best_id : (best_id | kTcUnsure);
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
                     float margin, int32_t* best, cudaStream_t stream) {
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
      tmap_x, tmap_c, low_dim, units, n_units, margin, best); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

// ---------------------------------------------------------------------------------------------
// Sparse-row variant (see kmeans_tc.cuh).  22 warps: warp 1 = MMA issuer (also loads the bucket's
// centroids with TMA when the bucket changes), warps 2-5 = epilogue, warps 6-21 = builders (four
// threads per 8-slot chunk of the tile's sparse block); warp 0 only sets up.  Shared memory: the whole A tile
// (num_kb x 16 KiB, the layout TMA's SWIZZLE_128B would produce: 128-byte rows, 16-byte chunk index
// XOR (row & 7), 8-row groups 1 KiB apart, one 16 KiB block per 64 columns) and the bucket's
// centroids (num_kb x 8 KiB: 64 rows per block).  A single A buffer: building tile k + 1 waits for
// the MMAs of tile k; the builders' row loads for tile k + 1 are in flight meanwhile (registers).
constexpr int kSpBuilders = 512;                          // 16 warps
constexpr int kSpThreads = 6 * 32 + kSpBuilders;
constexpr int kSpMaxKb = kSparseMaxDim / kBoxK;            // 7
constexpr int kSpABytes = kSpMaxKb * kBoxBytes;            // 112 KiB
constexpr int kSpBBlock = kSparseMaxLists * kBoxK * 2;     // 8 KiB
constexpr int kSpBBytes = kSpMaxKb * kSpBBlock;            // 56 KiB
constexpr int kSpSmemBytes = kSpABytes + kSpBBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kSpTmemCols = 128;                           // 2 accumulators x 64 columns
constexpr int kSpChunks = kSparseMaxWidth / 8;

// -0.0f also has the sign bit set but is not negative; an OR of bit patterns cannot tell, so any
// set sign bit counts (conservative: the general margin is used).
__device__ __forceinline__ bool vv_is_negative(uint32_t ored_bits) { return (ored_bits >> 31) != 0u; }

__global__ void __launch_bounds__(kSpThreads, 1)
kmeans_tc_sparse_kernel(const __grid_constant__ CUtensorMap tmap_c, const uint16_t* __restrict__ ell_idx,
                        const float* __restrict__ ell_val, const uint16_t* __restrict__ ell_nnz, int32_t W,
                        uint32_t low_dim, const int4* __restrict__ units, const int32_t* __restrict__ n_units_ptr,
                        float margin, int32_t* __restrict__ neg_seen, int use_neg_seen,
                        int32_t* __restrict__ best_out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_base = (raw + 1023u) & ~1023u;
  const uint32_t b_base = a_base + kSpABytes;
  const uint32_t bar_base = b_base + kSpBBytes;
  const uint32_t a_full = bar_base, a_empty = bar_base + 8u, b_full = bar_base + 16u;
  auto tfull_bar = [&](int a) { return bar_base + 24u + 8u * a; };
  auto tempty_bar = [&](int a) { return bar_base + 40u + 8u * a; };
  const uint32_t tmem_slot = bar_base + 56u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  unsigned char* a_ptr = smem_raw + (a_base - raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(a_full, kSpBuilders / 32);
    mbar_init(a_empty, 1);
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(kSpTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int num_kb = static_cast<int>((low_dim + kBoxK - 1) / kBoxK);
  const int64_t n_units = *n_units_ptr;
  const int64_t per = (n_units + gridDim.x - 1) / gridDim.x;
  const int64_t u0 = min(n_units, per * blockIdx.x), u1 = min(n_units, u0 + per);

  if (warp == 1) {
    // ===================== MMA issuer (+ centroid loads) =====================
    int acc = 0;
    uint32_t acc_phase = 0, b_phase = 0;
    int cur_c0 = -1;
    const uint64_t adesc0 = make_smem_desc(a_base), bdesc0 = make_smem_desc(b_base);
    for (int64_t u = u0; u < u1; ++u) {
      const uint32_t k = static_cast<uint32_t>(u - u0);
      const int4 d = __ldg(units + u);
      const int nc = min(kSparseMaxLists, d.w - d.z);
      const uint32_t n_mma = static_cast<uint32_t>(max(16, (nc + 15) & ~15));
      const uint32_t idesc = make_idesc(kTileM, n_mma);
      if (d.z != cur_c0) {
        // every earlier MMA has retired (commit of the previous tile): the centroid block is free
        if (k > 0) mbar_wait(a_empty, (k - 1u) & 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(num_kb) * kSpBBlock);
          for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(b_base + kb * kSpBBlock, &tmap_c, kb * kBoxK, d.z, b_full);
        }
        mbar_wait(b_full, b_phase);
        b_phase ^= 1u;
        cur_c0 = d.z;
      }
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      mbar_wait(a_full, k & 1u);
      tc_fence_after();
      if (lane == 0) {
        // the issue loop is on the tile's critical path: descriptors advance by adding to the start-address field
        // (units of 16 bytes; the operands sit below 256 KiB, so the field cannot overflow)
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kSparseMaxLists);
        uint64_t ad = adesc0, bd = bdesc0;
        int rem = static_cast<int>(low_dim);
        for (int kb = 0; kb < num_kb; ++kb, ad += kBoxBytes >> 4, bd += kSpBBlock >> 4, rem -= kBoxK) {
          if (rem >= kBoxK) {
            tc_mma_bf16(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
            tc_mma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
            tc_mma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
            tc_mma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
          } else {
            for (int ks = 0; ks < (rem + 15) / 16; ++ks)
              tc_mma_bf16(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc, (kb | ks) != 0 ? 1u : 0u);
          }
        }
        tc_commit(a_empty);
        tc_commit(tfull_bar(acc));
      }
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp >= 2 && warp < 6) {
    // ===================== epilogue: top-2 of every row =====================
    const int quarter = warp & 3;  // TMEM lanes [32 * quarter, 32 * quarter + 32)
    int acc = 0;
    uint32_t acc_phase = 0;
    constexpr float kRelErr = 0.005f;
    const bool tight = use_neg_seen != 0 && *neg_seen == 0;
    for (int64_t u = u0; u < u1; ++u) {
      const int4 d = __ldg(units + u);
      const int nc = min(kSparseMaxLists, d.w - d.z);
      const int q = d.x + quarter * 32 + lane;
      float best = -INFINITY, second = -INFINITY;
      int best_id = 0;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int chunks = (nc + 31) >> 5;
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(acc * kSparseMaxLists + ch * 32);
        tc_ld_32x32(taddr, v);
        const int cols = min(32, nc - ch * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = __uint_as_float(v[j]);
          if (j < cols) {
            if (s > best) { second = best; best = s; best_id = ch * 32 + j; }
            else if (s > second) { second = s; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      // Non-negative operands (what spectra are): every product is >= 0, so a score's bf16 error is at most
      // 2^-8 of the score itself (two roundings of 2^-9; plus the accumulator's own rounding), and the two
      // best lists are in the right order once they differ by more than kRelErr * (best + second) -- a few
      // times tighter than the general bound `margin`.  Whether any value was negative is known from the
      // previous launch (neg_seen).
      const float need = tight ? kRelErr * (best + second) + 1e-5f : margin;
      if (q < d.y) best_out[q] = (best - second > need) ? best_id : (best_id | kTcUnsure);
    }
  } else if (warp >= 6) {
    // ===================== builders: sparse rows -> swizzled bf16 tile =====================
    // The tile's sparse rows are one contiguous block of 128 * W / 8 chunks of 8 slots.  Builder
    // thread bt owns chunks bt, bt + 512, ... (the same rows and chunk positions for every tile), so a
    // warp's loads are contiguous (coalesced) and its lanes scatter different chunks of the same few
    // rows -- columns ascend along a row, so their banks differ.  Slots past a row's population are
    // loaded anyway (no dependent load on the population) and skipped in the scatter.
    const int bt = threadIdx.x - 6 * 32;  // 0..kSpBuilders - 1
    constexpr int kMine = kSpChunks * 128 / kSpBuilders;  // chunks per thread at the widest rows
    const int wc = W >> 3;                // chunks per row
    const int n_chunks = 128 * wc;
    int row_i[kMine], slot0_i[kMine];
    uint32_t base_i[kMine], x7_i[kMine];
#pragma unroll
    for (int i = 0; i < kMine; ++i) {
      const int g = bt + kSpBuilders * i;
      const int row = g / wc;
      row_i[i] = g < n_chunks ? row : 0x7fffffff;  // never below a tile's row count
      slot0_i[i] = (g - row * wc) * 8;
      base_i[i] = static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
      x7_i[i] = static_cast<uint32_t>(row & 7) << 4;
    }
    uint4 ri[kMine];
    float4 rv[2 * kMine];
    int mm[kMine];  // population of the chunk's row (0: nothing to scatter); <= W by construction of the rows
    // two-deep register pipeline: the descriptor of tile u + 2 and the chunks of tile u + 1 are in
    // flight while tile u is built
    int4 d_far = make_int4(0, 0, 0, 0);
    bool far_ok = false;
    auto load_desc = [&](int64_t u) {
      far_ok = u < u1;
      if (far_ok) d_far = __ldg(units + u);
    };
    auto load_rows = [&]() {  // chunks of the tile of the last load_desc
      const int nrows = far_ok ? min(128, d_far.y - d_far.x) : 0;
      const int64_t q0 = d_far.x;
      const uint4* ip = reinterpret_cast<const uint4*>(ell_idx + q0 * W);
      const float4* vp = reinterpret_cast<const float4*>(ell_val + q0 * W);
#pragma unroll
      for (int i = 0; i < kMine; ++i) {
        mm[i] = 0;
        if (row_i[i] < nrows) {
          const int g = bt + kSpBuilders * i;
          ri[i] = __ldg(ip + g);
          rv[2 * i] = __ldg(vp + 2 * g);
          rv[2 * i + 1] = __ldg(vp + 2 * g + 1);
          mm[i] = static_cast<int>(__ldg(ell_nnz + q0 + row_i[i]));  // not touched before the scatter: no stall here
        }
      }
    };
    load_desc(u0);
    load_rows();
    load_desc(u0 + 1);
    uint32_t sign_bits = 0u;  // OR of the scattered values' bit patterns: top bit = some value was negative
    for (int64_t u = u0; u < u1; ++u) {
      const uint32_t k = static_cast<uint32_t>(u - u0);
      if (k > 0) mbar_wait(a_empty, (k - 1u) & 1u);  // the MMAs of the previous tile have read A
      uint4* az = reinterpret_cast<uint4*>(a_ptr);
#pragma unroll 4
      for (int o = bt; o < num_kb * (kBoxBytes / 16); o += kSpBuilders) az[o] = make_uint4(0u, 0u, 0u, 0u);
      asm volatile("bar.sync 1, %0;" ::"n"(kSpBuilders) : "memory");
#pragma unroll
      for (int i = 0; i < kMine; ++i) {
        if (slot0_i[i] < mm[i]) {
          const uint32_t kk[8] = {ri[i].x & 0xffffu, ri[i].x >> 16, ri[i].y & 0xffffu, ri[i].y >> 16,
                                  ri[i].z & 0xffffu, ri[i].z >> 16, ri[i].w & 0xffffu, ri[i].w >> 16};
          const float vv[8] = {rv[2 * i].x, rv[2 * i].y, rv[2 * i].z, rv[2 * i].w,
                               rv[2 * i + 1].x, rv[2 * i + 1].y, rv[2 * i + 1].z, rv[2 * i + 1].w};
          unsigned char* row_ptr = a_ptr + base_i[i];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            if (slot0_i[i] + t < mm[i]) {
              const uint32_t col = kk[t];
              // block of 64 columns, 16-byte chunk (swizzled with the row), element
              const uint32_t off = ((col >> 6) << 14) + ((((col >> 3) & 7u) << 4) ^ x7_i[i]) + ((col & 7u) << 1);
              *reinterpret_cast<__nv_bfloat16*>(row_ptr + off) = __float2bfloat16_rn(vv[t]);
              sign_bits |= __float_as_uint(vv[t]);
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the MMA's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      load_rows();         // tile u + 1: in flight during the MMAs and the next zero fill
      load_desc(u + 2);
    }
    if (vv_is_negative(sign_bits)) *neg_seen = 1;  // benign race: every writer stores 1
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kSpTmemCols)
                 : "memory");
  }
}

int launch_kmeans_tc_sparse(const uint16_t* ell_idx, const float* ell_val, const uint16_t* ell_nnz, int32_t ell_width,
                            const uint16_t* c_bf16, int64_t ld_c, int64_t total_centroids, uint32_t low_dim,
                            const int4* units, const int32_t* n_units, float margin, int32_t* neg_seen,
                            int use_neg_seen, int32_t* best, cudaStream_t stream) {
  FLC_REQUIRE(kmeans_tc_sparse_ok(low_dim, ell_width), "shape not supported by the sparse tensor-core assignment");
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(c_bf16) & 15) == 0 && (ld_c % 8) == 0, "bf16 centroids must be 16-byte aligned");
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(ell_idx) & 15) == 0 && (reinterpret_cast<uintptr_t>(ell_val) & 15) == 0,
              "sparse rows must be 16-byte aligned");
  CUtensorMap tmap_c;
  FLC_TRY(make_bf16_tmap(&tmap_c, c_bf16, static_cast<uint64_t>(total_centroids), low_dim, ld_c, kSparseMaxLists));
  static bool attr_set = false;
  if (!attr_set) {
    FLC_CUDA(cudaFuncSetAttribute(kmeans_tc_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmemBytes));
    attr_set = true;
  }
  timed("kmeans_tc_sparse", stream, [&] { kmeans_tc_sparse_kernel<<<kNumSMs, kSpThreads, kSpSmemBytes, stream>>>(
      tmap_c, ell_idx, ell_val, ell_nnz, ell_width, low_dim, units, n_units, margin, neg_seen, use_neg_seen, best); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // namespace flc
