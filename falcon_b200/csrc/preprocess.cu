// Spectrum preprocessing -- SURVEY 8f row 1: the step immediately before the hot
// path (falcon/cluster/spectrum.py:73-169 on top of spectrum_utils 0.3.5, which is
// not in the image; its published semantics are restated in oracle/preprocess.py):
//   1. keep peaks with mz_min <= m/z <= mz_max;            valid? (>= min_peaks peaks
//   2. drop peaks within `tol` Da of the precursor m/z of    spanning >= min_mz_range,
//      every charge state z, z-1, .., 1;                     re-checked after each step,
//   3. drop peaks at or below min_intensity * base peak,     spectrum.py:27-52)
//      keep the max_peaks_used most intense of the rest;
//   4. scale intensities (root / log2(1 + x) / rank), divide by the L2 norm.
// Peaks of a spectrum must be ascending in m/z (what MsmsSpectrum guarantees).
//
// One warp per spectrum.  The intensity rank needed for the top-k cut is found by
// a 4-pass radix select on the float bits with a per-warp shared-memory histogram;
// ties at the cut go to the later peaks (stable ascending argsort).  Survivors are
// compacted to the front of the spectrum's own range; a prefix sum of the counts
// and one copy kernel produce the CSR arrays flc_vectorize consumes.
//
// HBM-bound: 8 bytes per input peak read (twice when the top-k cut applies) +
// 8 bytes per surviving peak written twice + 25 bytes per spectrum.
#include <cub/cub.cuh>

#include "common.cuh"

namespace flc {

constexpr int kPreWarps = 8;
constexpr double kProtonMass = 1.0072766;  // spectrum_utils remove_precursor_peak

struct PreParams {
  const float* mz;
  const float* intensity;
  const int64_t* indptr;
  int64_t n;
  const double* precursor_mz;
  const int32_t* charge;
  int32_t min_peaks;
  float min_mz_range;
  float mz_min, mz_max;  // NaN: no bound
  float remove_tol;      // < 0: keep the precursor peak
  float min_intensity;   // < 0: no intensity filter
  int32_t max_peaks;     // 0: no top-k cut
  int scaling;           // 0 none, 1 root, 2 log, 3 rank
  uint8_t* flag;         // [n_peaks] scratch
  float* tmp_mz;         // [n_peaks] survivors, compacted at the spectrum's own offset
  float* tmp_int;
  float* tmp_scaled;     // [n_peaks] scaled + normalised intensities of the survivors
  int32_t* count;        // [n + 1] survivors per spectrum (0 for invalid spectra)
  uint8_t* valid;        // [n]
};

__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(kPreWarps * 32)
preprocess_kernel(const PreParams P) {
  __shared__ uint32_t hist_s[kPreWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t* hist = hist_s[warp];
  const int64_t s = static_cast<int64_t>(blockIdx.x) * kPreWarps + warp;
  if (s >= P.n) return;
  const int64_t p0 = P.indptr[s], p1 = P.indptr[s + 1];
  auto fail = [&]() {
    if (lane == 0) { P.count[s] = 0; P.valid[s] = 0; }
  };
  auto is_valid = [&](int cnt, float lo, float hi) { return cnt >= P.min_peaks && cnt > 0 && hi - lo >= P.min_mz_range; };

  // ---- 1 + 2: m/z window, precursor peaks; statistics after each step
  const double pmz = P.precursor_mz[s];
  const int z = max(P.charge ? P.charge[s] : 1, 1);  // an unknown charge counts as 1 (spectrum.py:140-147)
  const double neutral = (pmz - kProtonMass) * z;
  int cnt_r = 0, cnt_p = 0;
  float lo_r = INFINITY, hi_r = -INFINITY, lo_p = INFINITY, hi_p = -INFINITY, imax = 0.f;
  for (int64_t p = p0 + lane; p < p1; p += 32) {
    const float m = __ldg(P.mz + p);
    const bool in_r = !(m < P.mz_min) && !(m > P.mz_max);  // NaN bounds compare false
    bool keep = in_r;
    if (keep && P.remove_tol >= 0.f) {
      for (int c = z; c >= 1; --c) {
        const double rm = neutral / c + kProtonMass;
        if (fabs(static_cast<double>(m) - rm) <= static_cast<double>(P.remove_tol)) keep = false;
      }
    }
    P.flag[p] = keep ? 1 : 0;
    if (in_r) { ++cnt_r; lo_r = fminf(lo_r, m); hi_r = fmaxf(hi_r, m); }
    if (keep) { ++cnt_p; lo_p = fminf(lo_p, m); hi_p = fmaxf(hi_p, m); imax = fmaxf(imax, __ldg(P.intensity + p)); }
  }
  cnt_r = warp_sum_i32(cnt_r); cnt_p = warp_sum_i32(cnt_p);
  lo_r = warp_min_f32(lo_r); hi_r = warp_max_f32(hi_r);
  lo_p = warp_min_f32(lo_p); hi_p = warp_max_f32(hi_p);
  imax = warp_max_f32(imax);
  if (!is_valid(cnt_r, lo_r, hi_r) || !is_valid(cnt_p, lo_p, hi_p)) { fail(); return; }
  __syncwarp();

  // ---- 3: base-peak threshold and top-k cut
  int n_keep = cnt_p;
  if (P.min_intensity >= 0.f || P.max_peaks > 0) {
    const float thr = (P.min_intensity >= 0.f ? P.min_intensity : 0.f) * imax;
    int n_gt = 0;
    for (int64_t p = p0 + lane; p < p1; p += 32)
      n_gt += (P.flag[p] && __ldg(P.intensity + p) > thr) ? 1 : 0;
    n_gt = warp_sum_i32(n_gt);
    const int kmax = P.max_peaks > 0 ? P.max_peaks : cnt_p;
    // nothing above the threshold: the sorted scan of the reference stops at the last (largest) peak
    const int want = n_gt > 0 ? min(kmax, n_gt) : 1;
    // radix select: the want-th largest intensity (bit pattern) among the kept peaks
    uint32_t prefix = 0, mask = 0;
    int remaining = want;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = lane; i < 256; i += 32) hist[i] = 0;
      __syncwarp();
      for (int64_t p = p0 + lane; p < p1; p += 32) {
        if (P.flag[p]) {
          const uint32_t b = __float_as_uint(fmaxf(__ldg(P.intensity + p), 0.f));
          if ((b & mask) == prefix) atomicAdd(hist + ((b >> shift) & 0xffu), 1u);
        }
      }
      __syncwarp();
      // the digit that holds the remaining-th largest: every lane sums 8 digits, a suffix
      // scan over the lanes finds the lane with the cut, that lane walks its 8 digits
      uint32_t mine = 0;
#pragma unroll
      for (int t = 0; t < 8; ++t) mine += hist[8 * lane + t];
      uint32_t suffix = mine;  // sum over lanes >= this one
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t dn = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += dn;
      }
      const uint32_t holds = __ballot_sync(0xffffffffu, suffix >= static_cast<uint32_t>(remaining));
      const int cut_lane = 31 - __clz(holds);  // lane 0 always qualifies: remaining <= total
      int digit = 8 * lane + 7;
      uint32_t acc = suffix - mine;  // peaks in the lanes above
      for (; digit > 8 * lane; --digit) {
        const uint32_t h = hist[digit];
        if (acc + h >= static_cast<uint32_t>(remaining)) break;
        acc += h;
      }
      digit = __shfl_sync(0xffffffffu, digit, cut_lane);
      acc = __shfl_sync(0xffffffffu, acc, cut_lane);
      remaining -= static_cast<int>(acc);
      prefix |= static_cast<uint32_t>(digit) << shift;
      mask |= 0xffu << shift;
      __syncwarp();
    }
    // peaks above `prefix` are in; of the ties, the `remaining` LAST ones (stable ascending argsort)
    int ties_after = 0;  // ties seen so far walking from the end of the spectrum
    const int64_t len = p1 - p0;
    for (int64_t base = ((len + 31) / 32 - 1) * 32; base >= 0; base -= 32) {
      const int64_t p = p0 + base + lane;
      bool f = false, tie = false;
      if (p < p1 && P.flag[p]) {
        const uint32_t b = __float_as_uint(fmaxf(__ldg(P.intensity + p), 0.f));
        f = b > prefix;
        tie = b == prefix;
      }
      const uint32_t tb = __ballot_sync(0xffffffffu, tie);
      if (tie) {
        const int later = ties_after + __popc(tb & ~(below | (1u << lane)));  // ties at higher positions
        f = later < remaining;
      }
      ties_after += __popc(tb);
      if (p < p1) P.flag[p] = f ? 1 : 0;
    }
    n_keep = want;
    __syncwarp();
    // validity of what is left
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t p = p0 + lane; p < p1; p += 32)
      if (P.flag[p]) { const float m = __ldg(P.mz + p); lo = fminf(lo, m); hi = fmaxf(hi, m); }
    lo = warp_min_f32(lo); hi = warp_max_f32(hi);
    if (!is_valid(n_keep, lo, hi)) { fail(); return; }
  }

  // ---- compaction to the front of the spectrum's own range
  int out = 0;
  for (int64_t base = 0; base < p1 - p0; base += 32) {
    const int64_t p = p0 + base + lane;
    const bool f = p < p1 && P.flag[p];
    const uint32_t fb = __ballot_sync(0xffffffffu, f);
    if (f) {
      const int64_t q = p0 + out + __popc(fb & below);
      P.tmp_mz[q] = __ldg(P.mz + p);
      P.tmp_int[q] = __ldg(P.intensity + p);
    }
    out += __popc(fb);
  }
  __syncwarp();
  // ---- 4: scaling + L2 norm (float64 accumulation, one rounding)
  double ss = 0.0;
  for (int i = lane; i < out; i += 32) {
    float v = P.tmp_int[p0 + i];
    if (P.scaling == 1) {
      v = sqrtf(v);
    } else if (P.scaling == 2) {
      v = static_cast<float>(log1p(static_cast<double>(v)) / 0.6931471805599453);
    } else if (P.scaling == 3) {
      // rank: max_rank - (number of kept peaks that sort after this one, ties by position)
      const uint32_t b = __float_as_uint(fmaxf(v, 0.f));
      int after = 0;
      for (int j = 0; j < out; ++j) {
        const uint32_t bj = __float_as_uint(fmaxf(P.tmp_int[p0 + j], 0.f));
        after += (bj > b || (bj == b && j > i)) ? 1 : 0;
      }
      v = static_cast<float>((P.max_peaks > 0 ? P.max_peaks : out) - after);
    }
    P.tmp_scaled[p0 + i] = v;
    ss = fma(static_cast<double>(v), static_cast<double>(v), ss);
  }
  ss = warp_sum_f64(ss);
  const float nrm = static_cast<float>(sqrt(ss));
  __syncwarp();
  for (int i = lane; i < out; i += 32) P.tmp_scaled[p0 + i] = P.tmp_scaled[p0 + i] / nrm;
  if (lane == 0) { P.count[s] = out; P.valid[s] = 1; }
}

// Survivors of spectrum s: tmp[indptr[s] .. + count[s]) -> out[out_indptr[s] ..).
__global__ void __launch_bounds__(256)
preprocess_copy_kernel(const float* __restrict__ tmp_mz, const float* __restrict__ tmp_scaled,
                       const int64_t* __restrict__ indptr, const int64_t* __restrict__ out_indptr, int64_t n,
                       float* __restrict__ out_mz, float* __restrict__ out_intensity) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (s >= n) return;
  const int64_t src = indptr[s], dst = out_indptr[s], cnt = out_indptr[s + 1] - dst;
  for (int64_t i = lane; i < cnt; i += 32) {
    out_mz[dst + i] = tmp_mz[src + i];
    out_intensity[dst + i] = tmp_scaled[src + i];
  }
}

struct PreLayout {
  uint8_t* flag;
  float* tmp_mz;
  float* tmp_int;
  float* tmp_scaled;
  int32_t* count;
  void* cub_tmp;
  size_t cub_bytes;
};

static void pre_layout(Workspace& ws, int64_t n, int64_t n_peaks, PreLayout& L) {
  const size_t np = static_cast<size_t>(n_peaks > 0 ? n_peaks : 1);
  L.flag = ws.take<uint8_t>(np);
  L.tmp_mz = ws.take<float>(np);
  L.tmp_int = ws.take<float>(np);
  L.tmp_scaled = ws.take<float>(np);
  L.count = ws.take<int32_t>(static_cast<size_t>(n + 1));
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int64_t*)nullptr, static_cast<int>(n + 1));
  L.cub_bytes = b;
  L.cub_tmp = ws.take<char>(b);
}

}  // namespace flc

extern "C" {

size_t flc_preprocess_workspace_bytes(int64_t n, int64_t n_peaks) {
  if (n <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::PreLayout L;
  flc::pre_layout(ws, n, n_peaks, L);
  return ws.used + 256;
}

int flc_preprocess(const float* mz, const float* intensity, const int64_t* indptr, int64_t n, int64_t n_peaks,
                   const double* precursor_mz, const int32_t* charge, int32_t min_peaks, float min_mz_range,
                   float mz_min, float mz_max, float remove_precursor_tol, float min_intensity,
                   int32_t max_peaks_used, int scaling, float* out_mz, float* out_intensity,
                   int64_t* out_indptr, uint8_t* valid, int64_t* n_out_peaks, void* workspace,
                   size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31) && n_peaks >= 0, "bad sizes");
  FLC_REQUIRE(scaling >= 0 && scaling <= 3, "Unknown intensity scaling");
  FLC_REQUIRE(max_peaks_used >= 0 && min_peaks >= 0, "max_peaks_used / min_peaks must be non-negative");
  FLC_REQUIRE(n_out_peaks != nullptr, "null n_out_peaks");
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    *n_out_peaks = 0;
    FLC_CUDA(cudaMemsetAsync(out_indptr, 0, sizeof(int64_t), stream));
    FLC_CUDA(cudaStreamSynchronize(stream));
    return FLC_OK;
  }
  FLC_REQUIRE(indptr && precursor_mz && out_indptr && valid, "null argument");
  Workspace ws(workspace, workspace_bytes);
  PreLayout L;
  pre_layout(ws, n, n_peaks, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "preprocess workspace too small: need %zu", ws.used);
  PreParams P;
  P.mz = mz; P.intensity = intensity; P.indptr = indptr; P.n = n; P.precursor_mz = precursor_mz; P.charge = charge;
  P.min_peaks = min_peaks; P.min_mz_range = min_mz_range; P.mz_min = mz_min; P.mz_max = mz_max;
  P.remove_tol = remove_precursor_tol; P.min_intensity = min_intensity; P.max_peaks = max_peaks_used;
  P.scaling = scaling; P.flag = L.flag; P.tmp_mz = L.tmp_mz; P.tmp_int = L.tmp_int; P.tmp_scaled = L.tmp_scaled;
  P.count = L.count; P.valid = valid;
  FLC_CUDA(cudaMemsetAsync(L.count + n, 0, sizeof(int32_t), stream));
  timed("preprocess", stream, [&] {
    preprocess_kernel<<<static_cast<unsigned>((n + kPreWarps - 1) / kPreWarps), kPreWarps * 32, 0, stream>>>(P); });
  FLC_LAUNCH_CHECK();
  size_t tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.count, out_indptr, static_cast<int>(n + 1), stream));
  count_launch(2);
  timed("preprocess_copy", stream, [&] {
    preprocess_copy_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, stream>>>(
        L.tmp_mz, L.tmp_scaled, indptr, out_indptr, n, out_mz, out_intensity); });
  FLC_LAUNCH_CHECK();
  int64_t total = 0;
  FLC_CUDA(cudaMemcpyAsync(&total, out_indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  FLC_CUDA(cudaStreamSynchronize(stream));
  *n_out_peaks = total;
  return FLC_OK;
}

}  // extern "C"
