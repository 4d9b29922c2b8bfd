// Stage 1: binning + MurmurHash3 feature hashing + L2 normalisation.
//
// One warp per spectrum.  Lanes load (m/z, intensity) pairs from the two CSR
// peak arrays with coalesced 128-byte requests, compute the mass bin in float64
// (falcon/cluster/spectrum.py:291 evaluates in float64: float32 m/z with
// float64 scalars -- SURVEY F5a), hash the int32 bin with MurmurHash3_x86_32 in
// registers, and accumulate intensities into a per-warp shared-memory row.
// Peaks that collide on a hashed column are added in peak order
// (__match_any_sync gives every lane its rank among same-column lanes) so the
// float32 sum is bit-identical to the sequential reference loop (A.1).  The
// squared norm is reduced in float64 with shuffles.  Spectra of up to 64 peaks
// (falcon keeps at most 50) take HALF a warp each -- two spectra per warp share the
// bookkeeping instructions -- and never sweep the dense row: the first lane of
// every distinct column owns it and emits its value to the float32 row
// (optional), the bfloat16 row (tcgen05 scan operand, staged in shared memory
// and copied out 16 bytes per lane) and the sparse ELL copy that k-means and
// the exact re-scoring read (distinct columns, peak order, zero padded).
//
// HBM-bound: algorithmic bytes per spectrum = 8 * peaks + 16 (indptr) + 4 (order)
// + 4 * low_dim [f32] + 2 * ld_bf16 [bf16] + 6 * ell_width + 2 [ELL].
#include <cuda_bf16.h>

#include "common.cuh"

namespace flc {

constexpr int kVecWarps = 8;

__global__ void hash_table_kernel(uint32_t vec_len, uint32_t low_dim, uint32_t seed,
                                  uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < vec_len) out[i] = murmur3_int32(i, seed) % low_dim;
}

struct VecParams {
  const float* mz;
  const float* intensity;
  const int64_t* indptr;
  const int32_t* order;
  const int32_t* dest;
  int64_t n;
  double min_mz, bin_size, inv_bin;
  uint32_t vec_len, low_dim, mod_magic, seed;
  int norm;
  uint32_t row_len;  // low_dim rounded up to 64
  float* out_f32;
  int64_t ld_f32;
  uint16_t* out_bf16;
  int64_t ld_bf16;
  int32_t* out_hash_idx;
  uint16_t* ell_idx;
  float* ell_val;
  uint16_t* ell_nnz;
  int32_t ell_width;
  int32_t* ell_overflow;
};

struct Peak {
  uint32_t col;
  float x;
  bool valid;
};

// Bin + hash of one peak (m/z, intensity) at peak position p; `have` = the lane holds a peak.
__device__ __forceinline__ Peak make_peak(const VecParams& P, bool have, float m, float x, int64_t p,
                                          double vec_len_d) {
  Peak k{0u, x, false};
  if (have) {
    // floor((m - min_mz) / bin_size) in float64.  The quotient is first taken as a
    // product with 1 / bin_size (relative error < 2^-51); only when that lands within
    // 1e-4 of an integer is the exact division needed to get the reference's floor.
    const double t = static_cast<double>(m) - P.min_mz;
    const double q = t * P.inv_bin;
    double b = floor(q);
    const double frac = q - b;
    if (frac < 1e-4 || frac > 1.0 - 1e-4) b = floor(t / P.bin_size);
    if (b >= 0.0 && b < vec_len_d) {
      const uint32_t h = murmur3_int32(static_cast<uint32_t>(static_cast<int32_t>(b)), P.seed);
      // h % low_dim with a precomputed reciprocal: the quotient estimate is short by at most one
      k.col = h - __umulhi(h, P.mod_magic) * P.low_dim;
      if (k.col >= P.low_dim) k.col -= P.low_dim;
      k.valid = true;
    }
    if (P.out_hash_idx) P.out_hash_idx[p] = k.valid ? static_cast<int32_t>(k.col) : -1;
  }
  return k;
}

__device__ __forceinline__ Peak load_peak(const VecParams& P, int64_t p, int64_t p1, double vec_len_d) {
  const bool have = p < p1;
  return make_peak(P, have, have ? __ldg(P.mz + p) : 0.f, have ? __ldg(P.intensity + p) : 0.f, p, vec_len_d);
}

// Software pipeline of the spectrum loop: where a spectrum's peaks are (two
// iterations ahead) and its first 64 peaks (one iteration ahead) are already in
// flight while the current spectrum is hashed.
struct SpecMeta {
  int64_t r, p0, p1;
};
struct SpecRaw {
  float m0, x0, m1, x1;
};
__device__ __forceinline__ SpecMeta load_meta(const VecParams& P, int64_t rr) {
  SpecMeta s{0, 0, 0};
  if (rr < P.n) {
    const int64_t src = P.order ? static_cast<int64_t>(__ldg(P.order + rr)) : rr;
    s.r = P.dest ? static_cast<int64_t>(__ldg(P.dest + rr)) : rr;
    s.p0 = __ldg(P.indptr + src);
    s.p1 = __ldg(P.indptr + src + 1);
  }
  return s;
}
__device__ __forceinline__ SpecRaw load_raw(const VecParams& P, const SpecMeta& s, int lane) {
  SpecRaw w{0.f, 0.f, 0.f, 0.f};
  const int64_t pa = s.p0 + lane, pb = pa + 32;
  if (pa < s.p1) { w.m0 = __ldg(P.mz + pa); w.x0 = __ldg(P.intensity + pa); }
  if (pb < s.p1) { w.m1 = __ldg(P.mz + pb); w.x1 = __ldg(P.intensity + pb); }
  return w;
}

// row[col] += x for the 32 peaks of one pass, colliding lanes in peak order.
// Returns true for the first lane of every distinct (valid) column.
__device__ __forceinline__ bool accumulate_pass(float* row, const Peak& k, int lane, uint32_t below) {
  const uint32_t key = k.valid ? k.col : (0x80000000u | lane);
  const uint32_t same = __match_any_sync(0xffffffffu, key);
  const int rank = __popc(same & below);
  if (__all_sync(0xffffffffu, (same & (same - 1u)) == 0u)) {
    if (k.valid) row[k.col] += k.x;  // no two lanes share a column
  } else {
    int rounds = __popc(same);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
    for (int t = 0; t < rounds; ++t) {
      if (k.valid && rank == t) row[k.col] += k.x;
      __syncwarp();
    }
  }
  __syncwarp();
  return k.valid && rank == 0;
}

// ---------------------------------------------------------------- wide spectra (> 64 peaks)
// Full warp on one spectrum: accumulate pass by pass, then sweep the dense row for the
// norm and every output.  `row` must be all zero on entry and is all zero on exit.
__device__ void vectorize_wide(const VecParams& P, float* row, int64_t r, int64_t p0, int64_t p1, int lane,
                               double vec_len_d) {
  const uint32_t below = (1u << lane) - 1u;
  float* dst_f = P.out_f32 ? P.out_f32 + r * P.ld_f32 : nullptr;
  uint16_t* dst_b = P.out_bf16 ? P.out_bf16 + r * P.ld_bf16 : nullptr;
  uint16_t* di = P.ell_idx ? P.ell_idx + r * P.ell_width : nullptr;
  float* dv = P.ell_idx ? P.ell_val + r * P.ell_width : nullptr;
  for (int64_t base = p0; base < p1; base += 32)
    accumulate_pass(row, load_peak(P, base + lane, p1, vec_len_d), lane, below);
  double scale = 1.0;
  if (P.norm) {
    double ss = 0.0;
    for (uint32_t i = lane; i < P.low_dim; i += 32) ss = fma(static_cast<double>(row[i]), static_cast<double>(row[i]), ss);
    ss = warp_sum_f64(ss);
    scale = ss > 0.0 ? rsqrt(ss) : 1.0;
  }
  int count = 0;
  for (uint32_t base = 0; base < P.row_len; base += 32) {
    const uint32_t i = base + lane;
    const float sv = static_cast<float>(static_cast<double>(row[i]) * scale);
    row[i] = 0.f;
    if (dst_f && i < P.low_dim) dst_f[i] = sv;
    if (dst_b && static_cast<int64_t>(i) < P.ld_bf16) dst_b[i] = __bfloat16_as_ushort(__float2bfloat16_rn(sv));
    if (di) {
      const bool nz = sv != 0.f;
      const uint32_t bal = __ballot_sync(0xffffffffu, nz);
      const int pos = count + __popc(bal & below);
      if (nz && pos < P.ell_width) { di[pos] = static_cast<uint16_t>(i); dv[pos] = sv; }
      count += __popc(bal);
    }
  }
  if (di) {
    for (int pos = count + lane; pos < P.ell_width; pos += 32) {  // zero padding
      di[pos] = 0;
      dv[pos] = 0.f;
    }
    if (lane == 0) {
      if (P.ell_nnz) P.ell_nnz[r] = static_cast<uint16_t>(min(count, P.ell_width));
      if (count > P.ell_width) atomicMax(P.ell_overflow, count);
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------- main kernel
// HALF a warp per spectrum (falcon keeps at most 50 peaks: four passes of 16 lanes), two
// spectra per warp side by side, so the per-spectrum bookkeeping instructions are shared.
// Per half-warp shared memory: the float32 accumulation row, a bfloat16 staging row (both
// all-zero between spectra) and a stamp row that tells which columns an earlier pass of the
// same spectrum already owns.  Spectra with more than 64 peaks are handed to the full warp.
constexpr int kPasses = 4;
struct HalfRaw {
  float m[kPasses], x[kPasses];
};
__device__ __forceinline__ HalfRaw load_half_raw(const VecParams& P, const SpecMeta& s, int hl) {
  HalfRaw w;
#pragma unroll
  for (int t = 0; t < kPasses; ++t) {
    const int64_t p = s.p0 + 16 * t + hl;
    const bool have = p < s.p1 && s.p1 - s.p0 <= 16 * kPasses;
    w.m[t] = have ? __ldg(P.mz + p) : 0.f;
    w.x[t] = have ? __ldg(P.intensity + p) : 0.f;
  }
  return w;
}

__global__ void __launch_bounds__(kVecWarps * 32, 4)
vectorize_kernel(const VecParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int h = lane >> 4, hl = lane & 15;
  const size_t per_half = static_cast<size_t>(P.row_len) * 7;  // f32 row + bf16 row + 8-bit stamps
  float* row = reinterpret_cast<float*>(smem_raw + (warp * 2 + h) * per_half);
  uint16_t* brow = reinterpret_cast<uint16_t*>(row + P.row_len);
  uint8_t* stamp = reinterpret_cast<uint8_t*>(brow + P.row_len);
  const int64_t halves_total = static_cast<int64_t>(gridDim.x) * kVecWarps * 2;
  const uint32_t below = (1u << lane) - 1u;
  const uint32_t below_h = (1u << hl) - 1u;
  const double vec_len_d = static_cast<double>(P.vec_len);
  const bool bf16_vec = P.out_bf16 != nullptr && (P.ld_bf16 & 7) == 0 && (reinterpret_cast<uintptr_t>(P.out_bf16) & 15) == 0;

  for (uint32_t i = hl; i < P.row_len; i += 16) {
    row[i] = 0.f;
    brow[i] = 0;
    stamp[i] = 0;
  }
  __syncwarp();
  uint32_t serial = 0;

  // software pipeline: offsets two spectra ahead, raw peaks one spectrum ahead
  const int64_t rr0 = (static_cast<int64_t>(blockIdx.x) * kVecWarps + warp) * 2 + h;
  SpecMeta cur = load_meta(P, rr0), nxt = load_meta(P, rr0 + halves_total);
  HalfRaw raw = load_half_raw(P, cur, hl);
  for (int64_t rr = rr0; __any_sync(0xffffffffu, rr < P.n); rr += halves_total) {
    const SpecMeta nn = load_meta(P, rr + 2 * halves_total);
    const HalfRaw raw_nxt = load_half_raw(P, nxt, hl);
    const bool active = rr < P.n;
    const int64_t r = cur.r, p0 = cur.p0, p1 = active ? cur.p1 : cur.p0;
    const HalfRaw w = raw;
    cur = nxt;
    nxt = nn;
    raw = raw_nxt;
    const int np = static_cast<int>(min(p1 - p0, static_cast<int64_t>(1 << 20)));

    if (__any_sync(0xffffffffu, np > 16 * kPasses)) {
      // a wide spectrum in this pair: the full warp takes both spectra one after the other
      for (int hh = 0; hh < 2; ++hh) {
        const int src_lane = hh * 16;
        const int64_t wr = __shfl_sync(0xffffffffu, r, src_lane);
        const int64_t wp0 = __shfl_sync(0xffffffffu, p0, src_lane);
        const int64_t wp1 = __shfl_sync(0xffffffffu, p1, src_lane);
        const int wact = __shfl_sync(0xffffffffu, active ? 1 : 0, src_lane);
        float* wrow = reinterpret_cast<float*>(smem_raw + (warp * 2 + hh) * per_half);
        if (wact) vectorize_wide(P, wrow, wr, wp0, wp1, lane, vec_len_d);
      }
      continue;
    }
    if (++serial == 0x100u) {  // stamps are 8 bit: start over before a tag could repeat
      for (uint32_t i = hl; i < P.row_len; i += 16) stamp[i] = 0;
      serial = 1;
      __syncwarp();
    }
    const uint8_t tag = static_cast<uint8_t>(serial);
    int np_max = max(__shfl_sync(0xffffffffu, np, 0), __shfl_sync(0xffffffffu, np, 16));

    uint32_t col[kPasses];
    float val[kPasses];
    bool own[kPasses];
#pragma unroll
    for (int t = 0; t < kPasses; ++t) {
      own[t] = false;
      col[t] = 0;
      val[t] = 0.f;
      if (16 * t < np_max) {  // uniform
        const int64_t p = p0 + 16 * t + hl;
        const Peak k = make_peak(P, p < p1, w.m[t], w.x[t], p, vec_len_d);
        col[t] = k.col;
        // rank among the lanes of this half that hit the same column (peak order)
        const uint32_t key = k.valid ? (k.col | (static_cast<uint32_t>(h) << 20)) : (0x80000000u | lane);
        const uint32_t same = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(same & below);
        if (__all_sync(0xffffffffu, (same & (same - 1u)) == 0u)) {
          if (k.valid) row[k.col] += k.x;  // no two lanes share a column
        } else {
          int rounds = __popc(same);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
          for (int q = 0; q < rounds; ++q) {
            if (k.valid && rank == q) row[k.col] += k.x;
            __syncwarp();
          }
        }
        // the first lane of a column owns it unless an earlier pass already does
        if (k.valid && rank == 0 && stamp[k.col] != tag) {
          own[t] = true;
          stamp[k.col] = tag;
        }
        __syncwarp();
      }
    }
    // ---- norm over the owned columns (float64, within the half)
    double scale = 1.0;
    {
      double ss = 0.0;
#pragma unroll
      for (int t = 0; t < kPasses; ++t) {
        if (own[t]) {
          val[t] = row[col[t]];
          row[col[t]] = 0.f;  // leave the row zeroed
          ss = fma(static_cast<double>(val[t]), static_cast<double>(val[t]), ss);
        }
      }
      if (P.norm) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        scale = ss > 0.0 ? rsqrt(ss) : 1.0;
      }
    }
#pragma unroll
    for (int t = 0; t < kPasses; ++t) val[t] = static_cast<float>(static_cast<double>(val[t]) * scale);

    // ---- outputs
    if (P.out_f32) {  // dense float32 row: zero fill, then the owners scatter
      float* dst_f = P.out_f32 + r * P.ld_f32;
      if (active)
        for (uint32_t i = hl; i < P.low_dim; i += 16) dst_f[i] = 0.f;
      __syncwarp();
#pragma unroll
      for (int t = 0; t < kPasses; ++t)
        if (own[t]) dst_f[col[t]] = val[t];
    }
    if (P.out_bf16) {
      uint16_t* dst_b = P.out_bf16 + r * P.ld_bf16;
#pragma unroll
      for (int t = 0; t < kPasses; ++t)
        if (own[t]) brow[col[t]] = __bfloat16_as_ushort(__float2bfloat16_rn(val[t]));
      __syncwarp();
      if (active) {
        if (bf16_vec) {
          for (int64_t i = 8 * hl; i < P.ld_bf16; i += 128) {
            *reinterpret_cast<uint4*>(dst_b + i) = *reinterpret_cast<const uint4*>(brow + i);
            *reinterpret_cast<uint4*>(brow + i) = make_uint4(0u, 0u, 0u, 0u);
          }
        } else {
          for (int64_t i = hl; i < P.ld_bf16; i += 16) {
            dst_b[i] = brow[i];
            brow[i] = 0;
          }
        }
      }
    }
    if (P.ell_idx) {  // sparse copy: the owners' non-zero values, in peak order
      uint16_t* di = P.ell_idx + r * P.ell_width;
      float* dv = P.ell_val + r * P.ell_width;
      int count = 0;
#pragma unroll
      for (int t = 0; t < kPasses; ++t) {
        if (16 * t < np_max) {  // uniform
          const bool nz = own[t] && val[t] != 0.f;
          const uint32_t bal = (__ballot_sync(0xffffffffu, nz) >> (16 * h)) & 0xffffu;
          const int pos = count + __popc(bal & below_h);
          if (nz && pos < P.ell_width) { di[pos] = static_cast<uint16_t>(col[t]); dv[pos] = val[t]; }
          count += __popc(bal);
        }
      }
      if (active) {
        for (int pos = count + hl; pos < P.ell_width; pos += 16) {  // zero padding
          di[pos] = 0;
          dv[pos] = 0.f;
        }
        if (hl == 0) {
          if (P.ell_nnz) P.ell_nnz[r] = static_cast<uint16_t>(min(count, P.ell_width));
          if (count > P.ell_width) atomicMax(P.ell_overflow, count);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace flc

extern "C" {

int flc_hash_table(uint32_t vec_len, uint32_t low_dim, uint32_t seed, uint32_t* out,
                   flc_stream_t stream) {
  FLC_REQUIRE(low_dim > 0, "low_dim must be positive");
  FLC_REQUIRE(out != nullptr || vec_len == 0, "null output");
  if (vec_len == 0) return FLC_OK;
  flc::timed("hash_table", stream, [&] { flc::hash_table_kernel<<<(vec_len + 255) / 256, 256, 0, flc::as_stream(stream)>>>(vec_len, low_dim,
                                                                                  seed, out); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

int flc_vectorize(const float* mz, const float* intensity, const int64_t* indptr,
                  const int32_t* order, const int32_t* dest, int64_t n, double min_mz, double bin_size,
                  uint32_t vec_len, uint32_t low_dim, uint32_t seed, int norm, float* out_f32,
                  int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16, int32_t* out_hash_idx,
                  uint16_t* ell_idx, float* ell_val, uint16_t* ell_nnz, int32_t ell_width,
                  int32_t* ell_overflow, flc_stream_t stream) {
  FLC_REQUIRE(n >= 0, "n must be non-negative");
  FLC_REQUIRE(low_dim > 0 && low_dim <= 8192, "low_dim must be in [1, 8192]");
  FLC_REQUIRE(bin_size > 0.0, "bin_size must be positive");
  FLC_REQUIRE(vec_len > 0, "vec_len must be positive");
  FLC_REQUIRE(!out_f32 || ld_f32 >= low_dim, "ld_f32 < low_dim");
  FLC_REQUIRE(!out_bf16 || ld_bf16 >= low_dim, "ld_bf16 < low_dim");
  FLC_REQUIRE((ell_idx == nullptr) == (ell_val == nullptr), "ell_idx and ell_val go together");
  FLC_REQUIRE(!ell_idx || (ell_width > 0 && ell_width <= 65535 && ell_overflow != nullptr && low_dim <= 65536),
              "ELL output needs 0 < ell_width < 65536, an overflow flag and low_dim <= 65536");
  FLC_REQUIRE(!ell_nnz || ell_idx, "ell_nnz needs the ELL arrays");
  if (n == 0) return FLC_OK;
  FLC_REQUIRE(indptr != nullptr, "null indptr");
  flc::VecParams P;
  P.mz = mz; P.intensity = intensity; P.indptr = indptr; P.order = order; P.dest = dest; P.n = n;
  P.min_mz = min_mz; P.bin_size = bin_size; P.inv_bin = 1.0 / bin_size;
  P.vec_len = vec_len; P.low_dim = low_dim; P.seed = seed; P.norm = norm;
  // floor((2^32 - 1) / low_dim): the quotient estimate __umulhi(h, magic) is short by at most one
  P.mod_magic = static_cast<uint32_t>(0xffffffffull / low_dim);
  P.row_len = (low_dim + 63u) & ~63u;
  P.out_f32 = out_f32; P.ld_f32 = ld_f32; P.out_bf16 = out_bf16; P.ld_bf16 = ld_bf16;
  P.out_hash_idx = out_hash_idx;
  P.ell_idx = ell_idx; P.ell_val = ell_val; P.ell_nnz = ell_nnz; P.ell_width = ell_width;
  P.ell_overflow = ell_overflow;
  FLC_REQUIRE(!out_bf16 || ld_bf16 <= static_cast<int64_t>(P.row_len), "ld_bf16 exceeds low_dim rounded up to 64");
  const size_t smem = static_cast<size_t>(flc::kVecWarps) * 2 * P.row_len * 7;  // per half-warp: f32 row + bf16 row + stamps
  FLC_REQUIRE(smem <= 200 * 1024, "low_dim too large");
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(flc::vectorize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  int64_t blocks = (n + 2 * flc::kVecWarps - 1) / (2 * flc::kVecWarps);
  const int64_t max_blocks = static_cast<int64_t>(flc::kNumSMs) * 4;  // what fits an SM at once (shared memory, registers)
  if (blocks > max_blocks) blocks = max_blocks;
  flc::timed("vectorize", stream, [&] {
    flc::vectorize_kernel<<<static_cast<unsigned>(blocks), flc::kVecWarps * 32, smem, flc::as_stream(stream)>>>(P); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
