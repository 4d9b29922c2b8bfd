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
// squared norm is reduced in float64 with warp shuffles; the row leaves as
// float32 (exact re-scoring, k-means) and bfloat16 (tcgen05 scan operand).
//
// HBM-bound: algorithmic bytes per spectrum = 8 * peaks + 16 (indptr) +
// 4 * low_dim (+ 2 * ld_bf16 for the bf16 copy).
#include "common.cuh"

namespace flc {

constexpr int kVecWarps = 8;

__global__ void hash_table_kernel(uint32_t vec_len, uint32_t low_dim, uint32_t seed,
                                  uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < vec_len) out[i] = murmur3_int32(i, seed) % low_dim;
}

__global__ void __launch_bounds__(kVecWarps * 32)
vectorize_kernel(const float* __restrict__ mz, const float* __restrict__ intensity,
                 const int64_t* __restrict__ indptr, const int32_t* __restrict__ order,
                 int64_t n, double min_mz, double bin_size, uint32_t vec_len,
                 uint32_t low_dim, uint32_t seed, int norm,
                 float* __restrict__ out_f32, int64_t ld_f32,
                 uint16_t* __restrict__ out_bf16, int64_t ld_bf16,
                 int32_t* __restrict__ out_hash_idx,
                 uint16_t* __restrict__ ell_idx, float* __restrict__ ell_val, uint16_t* __restrict__ ell_nnz,
                 int32_t ell_width, int32_t* __restrict__ ell_overflow) {
  extern __shared__ float smem_rows[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* row = smem_rows + static_cast<size_t>(warp) * low_dim;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kVecWarps;

  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kVecWarps + warp; r < n; r += warps_total) {
    const int64_t src = order ? static_cast<int64_t>(order[r]) : r;
    const int64_t p0 = indptr[src];
    const int64_t p1 = indptr[src + 1];
    for (uint32_t i = lane; i < low_dim; i += 32) row[i] = 0.f;
    __syncwarp();

    for (int64_t base = p0; base < p1; base += 32) {
      const int64_t p = base + lane;
      const bool in_range = p < p1;
      float x = 0.f;
      uint32_t col = 0;
      bool valid = false;
      if (in_range) {
        const float m = __ldg(mz + p);
        x = __ldg(intensity + p);
        const double b = floor((static_cast<double>(m) - min_mz) / bin_size);
        if (b >= 0.0 && b < static_cast<double>(vec_len)) {
          col = murmur3_int32(static_cast<uint32_t>(static_cast<int32_t>(b)), seed) % low_dim;
          valid = true;
        }
        if (out_hash_idx) out_hash_idx[p] = valid ? static_cast<int32_t>(col) : -1;
      }
      // Rank of this lane among the lanes that hit the same column (peak order).
      const uint32_t key = valid ? col : (0x80000000u | lane);
      const uint32_t same = __match_any_sync(0xffffffffu, key);
      const int rank = __popc(same & ((1u << lane) - 1u));
      int rounds = __popc(same);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
      for (int t = 0; t < rounds; ++t) {
        if (valid && rank == t) row[col] += x;
        __syncwarp();
      }
    }

    double scale = 1.0;
    if (norm) {
      double ss = 0.0;
      for (uint32_t i = lane; i < low_dim; i += 32) {
        const double v = static_cast<double>(row[i]);
        ss = fma(v, v, ss);
      }
      ss = warp_sum_f64(ss);
      scale = ss > 0.0 ? 1.0 / sqrt(ss) : 1.0;
      // normalise in place: every output below reads the final float32 value
      for (uint32_t i = lane; i < low_dim; i += 32)
        row[i] = static_cast<float>(static_cast<double>(row[i]) * scale);
      __syncwarp();
    }
    if (out_f32) {
      float* dst = out_f32 + r * ld_f32;
      if (((low_dim | ld_f32) & 3) == 0) {
        for (uint32_t i = lane * 4; i < low_dim; i += 128)
          *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(row + i);
      } else {
        for (uint32_t i = lane; i < low_dim; i += 32) dst[i] = row[i];
      }
    }
    if (out_bf16) {
      uint16_t* dst = out_bf16 + r * ld_bf16;
      if ((ld_bf16 & 3) == 0) {
        for (int64_t i = lane * 4; i < ld_bf16; i += 128) {
          uint16_t h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t c = i + j;
            h[j] = f32_to_bf16_rne(c < low_dim ? row[c] : 0.f);
          }
          uint2 packed;
          packed.x = static_cast<uint32_t>(h[0]) | (static_cast<uint32_t>(h[1]) << 16);
          packed.y = static_cast<uint32_t>(h[2]) | (static_cast<uint32_t>(h[3]) << 16);
          *reinterpret_cast<uint2*>(dst + i) = packed;
        }
      } else {
        for (int64_t i = lane; i < ld_bf16; i += 32) dst[i] = f32_to_bf16_rne(i < low_dim ? row[i] : 0.f);
      }
    }
    if (ell_idx) {
      // Sparse (ELL) copy: the non-zero columns in ascending order, zero padded.
      uint16_t* di = ell_idx + r * ell_width;
      float* dv = ell_val + r * ell_width;
      int count = 0;
      for (uint32_t base = 0; base < low_dim; base += 32) {
        const uint32_t i = base + lane;
        const float v = i < low_dim ? row[i] : 0.f;
        const bool nz = v != 0.f;
        const uint32_t ballot = __ballot_sync(0xffffffffu, nz);
        const int pos = count + __popc(ballot & ((1u << lane) - 1u));
        if (nz && pos < ell_width) {
          di[pos] = static_cast<uint16_t>(i);
          dv[pos] = v;
        }
        count += __popc(ballot);
      }
      for (int pos = count + lane; pos < ell_width; pos += 32) {
        di[pos] = 0;
        dv[pos] = 0.f;
      }
      if (ell_nnz && lane == 0) ell_nnz[r] = static_cast<uint16_t>(min(count, ell_width));
      if (count > ell_width && lane == 0) atomicMax(ell_overflow, count);
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
                  const int32_t* order, int64_t n, double min_mz, double bin_size,
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
  const size_t smem = static_cast<size_t>(flc::kVecWarps) * low_dim * sizeof(float);
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(flc::vectorize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  int64_t blocks = (n + flc::kVecWarps - 1) / flc::kVecWarps;
  const int64_t max_blocks = static_cast<int64_t>(flc::kNumSMs) * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  flc::timed("vectorize", stream, [&] { flc::vectorize_kernel<<<static_cast<unsigned>(blocks), flc::kVecWarps * 32, smem,
                          flc::as_stream(stream)>>>(mz, intensity, indptr, order, n, min_mz, bin_size,
                                                    vec_len, low_dim, seed, norm, out_f32, ld_f32,
                                                    out_bf16, ld_bf16, out_hash_idx, ell_idx, ell_val,
                                                    ell_nnz, ell_width, ell_overflow); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
