// Stage 3a: inverted-list inner-product scan on tcgen05 tensor cores.
//
// For every precursor bucket the kernel forms S = X_b * X_b^T tile by tile
// (128 queries x up to 256 candidates, K = low_dim) and emits the (query,
// candidate) pairs whose bf16 inner product clears the threshold
// `1 - eps - margin`; the exact re-score, top-k and filter run on the few
// survivors (refine.cu).  This is the dense contraction of faiss'
// IndexIVFFlat/IndexFlatIP scanner restated for Blackwell:
//
//   warp 0  TMA producer: one elected lane streams 128x64 bf16 boxes of the
//           query rows (A) and candidate rows (B) through a 4-stage shared-memory
//           ring (cp.async.bulk.tensor, SWIZZLE_128B, completion on mbarriers).
//           Rows past the end of the matrix and columns past low_dim are
//           zero-filled by TMA, so low_dim = 400 needs no padded copy.
//   warp 1  MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16
//           (M = 128, N = candidates rounded up to 16, K = 16 per instruction)
//           from shared-memory descriptors into one of two 256-column fp32
//           accumulators in TMEM; tcgen05.commit releases ring slots / publishes
//           the accumulator.
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), threshold
//           compare into a per-row bit mask, warp prefix sum, ONE atomicAdd per
//           warp and 32-column chunk that has survivors, coalesced pair stores.
//
// S is symmetric: a query tile only meets the candidate tiles at or above its diagonal, and the
// epilogue emits (q, c) and (c, q) for every hit with c > q -- about half the MMAs of a large bucket.
//
// Persistent grid: one CTA per SM.  Query tiles are dealt round-robin, so the CTAs
// running together stream the candidate rows of the same bucket(s): HBM sees a
// row once, L2 serves the re-reads.
//
// Roofline: tensor-core bound for buckets >~ 4k rows, 2 * low_dim * n_b^2 FLOP
// per bucket; HBM/L2-latency bound for small buckets (n_b * low_dim * 2 bytes
// read once).
#include "tc_common.cuh"

namespace flc {

// ---------------------------------------------------------------- tile walker
struct Tile {
  int64_t q0;   // first query row
  int64_t c0;   // first candidate row
  int64_t end;  // end row of the bucket
};

// Walks the tiles of one CTA.  A unit = one query tile (128 rows) against the
// candidate tiles of its bucket from its own first row on.  Units are dealt
// round-robin in PAIRS of consecutive descriptor positions (pair p goes to CTA
// p mod grid; unit_desc_kernel orders a bucket's units so that every pair costs
// the same): the CTAs running at any moment stream the candidate rows of the same
// one or two buckets, so those rows come from L2, and the CTAs finish together.
struct TileWalker {
  const int4* unit_desc;
  int64_t total, stride, u;  // u: even position of the current pair
  int4 cur, nxt;  // descriptor of the current position and (prefetched) of the next one
  int ct, half;

  __device__ void init(const int4* ud, int64_t n_units, int64_t first, int64_t step) {
    unit_desc = ud; total = n_units; stride = 2 * step; u = 2 * first; ct = 0; half = 0;
    cur = nxt = make_int4(0, 0, 0, 0);
    if (u < total) cur = __ldg(unit_desc + u);
    if (u + 1 < total) nxt = __ldg(unit_desc + u + 1);
  }
  __device__ bool valid() const { return u + half < total; }
  __device__ Tile get() const {
    Tile t;
    t.q0 = cur.x;
    // candidate tiles from the bucket's end back to the query tile's own rows: the units of a bucket that
    // run together then read the same tile at the same time (two interleaved families, by the parity of
    // the query tile), so a row comes from HBM once or twice and from L2 for everybody else
    t.c0 = static_cast<int64_t>(cur.y) + static_cast<int64_t>(cur.w - 1 - ct) * kTileN;
    t.end = cur.z;
    return t;
  }
  __device__ void next() {
    if (++ct >= cur.w) {
      ct = 0;
      if (half == 0) { half = 1; } else { half = 0; u += stride; }
      cur = nxt;
      const int64_t p = half == 0 ? u + 1 : u + stride;  // the position after the new current one
      if (p < total) nxt = __ldg(unit_desc + p);
    }
  }
};

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kScanThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t low_dim,
               const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
               const int64_t* __restrict__ tile_off, const int4* __restrict__ unit_desc, float threshold,
               uint64_t* __restrict__ pairs, uint64_t capacity, unsigned long long* __restrict__ pair_count) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles_base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-byte alignment
  const uint32_t bar_base = tiles_base + kStages * kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM pointer
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

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
  const int64_t n_units = tile_off[n_buckets];
  (void)bucket_ptr;

  {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        TileWalker w;
        int stage = 0;
        uint32_t phase = 0;
        for (w.init(unit_desc, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
          const Tile tile = w.get();
          const bool two = (tile.end - tile.c0) > kBoxRows;
          const uint32_t bytes = kABytes + (two ? 2 : 1) * kBoxBytes;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), bytes);
            const uint32_t sa = tiles_base + stage * kStageBytes;
            const uint32_t sb = sa + kABytes;
            tma_load_2d(sa, &tmap, kb * kBoxK, static_cast<int32_t>(tile.q0), full_bar(stage));
            tma_load_2d(sb, &tmap, kb * kBoxK, static_cast<int32_t>(tile.c0), full_bar(stage));
            if (two)
              tma_load_2d(sb + kBoxBytes, &tmap, kb * kBoxK, static_cast<int32_t>(tile.c0 + kBoxRows),
                          full_bar(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      TileWalker w;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (w.init(unit_desc, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
        const Tile tile = w.get();
        const int64_t nc = min(static_cast<int64_t>(kTileN), tile.end - tile.c0);
        const uint32_t n_mma = static_cast<uint32_t>(max(static_cast<int64_t>(16), (nc + 15) & ~int64_t(15)));
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
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * 32);
              const uint64_t bdesc = make_smem_desc(sb + k * 32);
              tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
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
      // ===================== epilogue (warps 2..5) =====================
      const int quarter = warp & 3;  // TMEM lanes [32 * quarter, 32 * quarter + 32)
      TileWalker w;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (w.init(unit_desc, n_units, blockIdx.x, gridDim.x); w.valid(); w.next()) {
        const Tile tile = w.get();
        const int64_t nc = min(static_cast<int64_t>(kTileN), tile.end - tile.c0);
        const int64_t q = tile.q0 + quarter * 32 + lane;
        const bool q_valid = q < tile.end;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const int chunks = static_cast<int>((nc + 31) >> 5);
        for (int ch = 0; ch < chunks; ++ch) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(acc * kTileN + ch * 32);
          tc_ld_32x32(taddr, v);
          uint32_t mask = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) >= threshold ? 1u : 0u) << j;
          const int64_t cols_left = nc - ch * 32;
          if (cols_left < 32) mask &= (1u << cols_left) - 1u;
          if (!q_valid) mask = 0;
          // symmetry: only candidates at or above the diagonal count here (c >= q); every hit above it is
          // emitted in both directions, the tile below the diagonal is never computed
          const int64_t cbase_i = tile.c0 + ch * 32;
          const int64_t below = q - cbase_i;  // columns j < below have c < q
          if (below >= 32) mask = 0;
          else if (below > 0) mask &= ~((1u << below) - 1u);
          const uint32_t diag = (below >= 0 && below < 32) ? (mask & (1u << below)) : 0u;
          const uint32_t cnt = 2u * __popc(mask) - (diag != 0u ? 1u : 0u);
          uint32_t incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
          }
          const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
          if (warp_total != 0) {
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(pair_count, static_cast<unsigned long long>(warp_total));
            base = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long pos = base + incl - cnt;
            const uint64_t qq = static_cast<uint64_t>(q);
            const uint64_t cbase = static_cast<uint64_t>(cbase_i);
            while (mask) {
              const int j = __ffs(mask) - 1;
              mask &= mask - 1;
              const uint64_t c = cbase + j;
              if (pos < capacity) pairs[pos] = (qq << 32) | c;
              ++pos;
              if (c != qq) {
                if (pos < capacity) pairs[pos] = (c << 32) | qq;
                ++pos;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
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

int launch_scan_tc(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, uint32_t low_dim,
                   const int64_t* bucket_ptr, int64_t n_buckets, const int64_t* tile_off,
                   const int4* unit_desc, float threshold, uint64_t* pairs, uint64_t pair_capacity,
                   unsigned long long* pair_count, cudaStream_t stream) {
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, "x_bf16 must be 16-byte aligned");
  CUtensorMap tmap;
  FLC_TRY(make_bf16_tmap(&tmap, x_bf16, static_cast<uint64_t>(n), low_dim, ld_bf16));
  static bool attr_set = false;
  if (!attr_set) {
    FLC_CUDA(cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  timed("scan_tc", stream, [&] { scan_tc_kernel<<<kNumSMs, kScanThreads, kSmemBytes, stream>>>(tmap, low_dim, bucket_ptr, n_buckets, tile_off,
                                                               unit_desc, threshold, pairs, pair_capacity, pair_count); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // namespace flc
