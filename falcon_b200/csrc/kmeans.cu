// Stage 2: per-bucket inverted-file index -- sizing, spherical k-means training,
// coarse assignment and probe selection (faiss IndexIVFFlat over an
// IndexFlatIP quantiser, METRIC_INNER_PRODUCT).
//
// All buckets are trained together ("batched"): rows find their bucket by
// binary search in bucket_ptr, centroids find theirs in centroid_ptr, so one
// launch covers thousands of small buckets and a 100k-row bucket alike.
//   assign  : one warp per row, float32 inner products against the bucket's
//             centroids (L2 resident), arg-max with ties to the lower id.
//   update  : one CTA per centroid, threads own dimensions and walk the
//             bucket's rows in index order accumulating in float64 --
//             deterministic (no float atomics) and identical to the oracle's
//             sequential float64 sum.
//   fix     : one CTA per bucket: empty lists are re-seeded from the largest
//             list with the +-1/1024 perturbation faiss uses, then every
//             centroid is L2-normalised (spherical k-means for the IP metric).
// The final assignment / probe selection (flc_ivf_assign) uses float64 inner
// products so that it agrees with the oracle wherever there is no exact tie.
#include <algorithm>

#include "common.cuh"

namespace flc {

__host__ __device__ inline int32_t nlist_rule(int64_t n) {
  if (n < 100) return 0;
  if (n < 1000000) {
    int64_t p = 1;
    while (39 * p * 2 <= n) p *= 2;  // 2^floor(log2(n / 39))
    return static_cast<int32_t>(p);
  }
  if (n < 10000000) return 1 << 16;
  if (n < 100000000) return 1 << 18;
  return 1 << 20;
}

// Shared-memory budget of one bucket in the fused kernel: the ELL rows (float32
// values with an odd pitch, uint16 columns), the centroids ([column][list]
// float32), their float64 accumulators ([list][column]), counts, assignments.
__host__ __device__ inline size_t fused_smem_bytes(int64_t nb, int32_t L, int32_t W, uint32_t low_dim) {
  const size_t lp = static_cast<size_t>((L + 3) & ~3);
  size_t b = static_cast<size_t>(nb) * (W + 1) * 4;       // values
  b += static_cast<size_t>(nb) * (W + 2) * 2;             // columns
  b = (b + 15) & ~size_t(15);
  b += static_cast<size_t>(L) * low_dim * 8;              // accumulators
  b += static_cast<size_t>(low_dim) * lp * 4;             // centroids
  b += static_cast<size_t>(L) * 8 + ((static_cast<size_t>(nb) + 15) & ~size_t(15)) + 64;
  return b;
}
constexpr size_t kFusedSmemCap = 200 * 1024;   // largest bucket the fused trainer takes
constexpr size_t kFusedSmemSmall = 100 * 1024; // buckets below this run two CTAs per SM

__host__ __device__ inline bool bucket_is_fused(int64_t nb, int32_t L, int32_t W, uint32_t low_dim,
                                                size_t smem_limit) {
  return L > 0 && W > 0 && fused_smem_bytes(nb, L, W, low_dim) <= smem_limit;
}

// Single CTA: nlist / nprobe per bucket, exclusive scan -> centroid_ptr, max nprobe.
__global__ void __launch_bounds__(1024)
ivf_plan_kernel(const int64_t* __restrict__ bucket_ptr, int64_t n_buckets, int32_t n_probe, int exhaustive,
                int32_t* __restrict__ nlist, int32_t* __restrict__ nprobe, int64_t* __restrict__ centroid_ptr) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  __shared__ int32_t max_s;
  __shared__ unsigned long long max_nb_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_s = 0; max_s = 0; max_nb_s = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n_buckets; base += 1024) {
    const int64_t b = base + tid;
    int32_t L = 0, P = 0;
    if (b < n_buckets) {
      const int64_t nb = bucket_ptr[b + 1] - bucket_ptr[b];
      L = nlist_rule(nb);
      if (L > 0) {
        P = exhaustive ? L : max(1, min((L + 7) / 8, n_probe));
        atomicMax(&max_nb_s, static_cast<unsigned long long>(nb));
      }
      nlist[b] = L;
      nprobe[b] = P;
    }
    int64_t incl = L;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    int32_t pmax = P;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
    if (lane == 0) atomicMax(&max_s, pmax);
    __syncthreads();
    int64_t warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_sums[w];
    const int64_t carry = carry_s;
    if (b < n_buckets) centroid_ptr[b] = carry + warp_off + incl - L;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_off + incl;
    __syncthreads();
  }
  if (tid == 0) {
    centroid_ptr[n_buckets] = carry_s;
    centroid_ptr[n_buckets + 1] = max_s;
    centroid_ptr[n_buckets + 2] = static_cast<int64_t>(max_nb_s);
  }
}

__device__ __forceinline__ int64_t find_segment(const int64_t* __restrict__ ptr, int64_t n_seg, int64_t i) {
  int64_t lo = 0, hi = n_seg;  // last s with ptr[s] <= i
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void kmeans_init_kernel(const float* __restrict__ x, int64_t ld, uint32_t low_dim,
                                   const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                                   const int32_t* __restrict__ nlist, const int64_t* __restrict__ centroid_ptr,
                                   int64_t total, int32_t W, size_t fused_limit,
                                   float* __restrict__ centroids) {
  const int64_t gc = blockIdx.x;
  if (gc >= total) return;
  const int64_t b = find_segment(centroid_ptr, n_buckets, gc);
  const int64_t c = gc - centroid_ptr[b];
  const int64_t s = bucket_ptr[b], nb = bucket_ptr[b + 1] - s;
  if (bucket_is_fused(nb, nlist[b], W, low_dim, fused_limit)) return;
  const int64_t row = s + (c * nb) / nlist[b];
  for (uint32_t i = threadIdx.x; i < low_dim; i += blockDim.x)
    centroids[gc * low_dim + i] = x[row * ld + i];
}

__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ x, int64_t ld, int64_t n, uint32_t low_dim,
                     const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                     const int32_t* __restrict__ nlist, const int64_t* __restrict__ centroid_ptr,
                     const float* __restrict__ centroids, int32_t W, size_t fused_limit,
                     int32_t* __restrict__ assign) {
  extern __shared__ float smem_x[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (i >= n) return;
  const int64_t b = find_segment(bucket_ptr, n_buckets, i);
  const int32_t L = nlist[b];
  if (L == 0 || bucket_is_fused(bucket_ptr[b + 1] - bucket_ptr[b], L, W, low_dim, fused_limit)) {
    if (lane == 0) assign[i] = 0;
    return;
  }
  float* xi = smem_x + static_cast<size_t>(warp) * low_dim;
  for (uint32_t t = lane; t < low_dim; t += 32) xi[t] = x[i * ld + t];
  __syncwarp();
  const float* cent = centroids + centroid_ptr[b] * low_dim;
  float best = -INFINITY;
  int32_t best_c = 0;
  for (int32_t c = 0; c < L; ++c) {
    const float* cr = cent + static_cast<int64_t>(c) * low_dim;
    float acc = 0.f;
    for (uint32_t t = lane; t < low_dim; t += 32) acc = fmaf(xi[t], __ldg(cr + t), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (acc > best) { best = acc; best_c = c; }
  }
  if (lane == 0) assign[i] = best_c;
}

// One CTA per centroid: float64 sum of the rows assigned to it, in row order.
__global__ void __launch_bounds__(128)
kmeans_update_kernel(const float* __restrict__ x, int64_t ld, uint32_t low_dim,
                     const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                     const int64_t* __restrict__ centroid_ptr, int64_t total,
                     const int32_t* __restrict__ assign, const float* __restrict__ centroids,
                     const int32_t* __restrict__ nlist, int32_t W, size_t fused_limit,
                     float* __restrict__ new_centroids, double* __restrict__ counts) {
  const int64_t gc = blockIdx.x;
  if (gc >= total) return;
  const int64_t b = find_segment(centroid_ptr, n_buckets, gc);
  const int32_t c = static_cast<int32_t>(gc - centroid_ptr[b]);
  const int64_t s = bucket_ptr[b], e = bucket_ptr[b + 1];
  if (bucket_is_fused(e - s, nlist[b], W, low_dim, fused_limit)) return;
  constexpr int kMaxPerThread = 8;  // low_dim <= 128 * 8 handled in registers per pass
  for (uint32_t d0 = 0; d0 < low_dim; d0 += 128 * kMaxPerThread) {
    double acc[kMaxPerThread];
#pragma unroll
    for (int k = 0; k < kMaxPerThread; ++k) acc[k] = 0.0;
    int64_t cnt = 0;
    for (int64_t i = s; i < e; ++i) {
      if (__ldg(assign + i) == c) {
        ++cnt;
#pragma unroll
        for (int k = 0; k < kMaxPerThread; ++k) {
          const uint32_t t = d0 + threadIdx.x + 128 * k;
          if (t < low_dim) acc[k] += static_cast<double>(__ldg(x + i * ld + t));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxPerThread; ++k) {
      const uint32_t t = d0 + threadIdx.x + 128 * k;
      if (t < low_dim)
        new_centroids[gc * low_dim + t] =
            cnt > 0 ? static_cast<float>(acc[k] / static_cast<double>(cnt)) : centroids[gc * low_dim + t];
    }
    if (threadIdx.x == 0 && d0 == 0) counts[gc] = static_cast<double>(cnt);
  }
}

// One CTA per bucket: split the largest list into every empty one, normalise.
__global__ void __launch_bounds__(128)
kmeans_fix_kernel(uint32_t low_dim, int64_t n_buckets, const int32_t* __restrict__ nlist,
                  const int64_t* __restrict__ centroid_ptr, const int64_t* __restrict__ bucket_ptr, int32_t W,
                  size_t fused_limit, float* __restrict__ new_centroids,
                  double* __restrict__ counts, float* __restrict__ centroids) {
  const int64_t b = blockIdx.x;
  if (b >= n_buckets) return;
  const int32_t L = nlist[b];
  if (L == 0 || bucket_is_fused(bucket_ptr[b + 1] - bucket_ptr[b], L, W, low_dim, fused_limit)) return;
  const int64_t c0 = centroid_ptr[b];
  __shared__ int32_t cj_s;
  __shared__ double red[128];
  const float eps = 1.0f / 1024.0f;
  for (int32_t ci = 0; ci < L; ++ci) {
    if (counts[c0 + ci] > 0.0) continue;  // uniform: all threads read the same value
    __syncthreads();
    if (threadIdx.x == 0) {
      int32_t best = 0;
      double bc = counts[c0];
      for (int32_t c = 1; c < L; ++c)
        if (counts[c0 + c] > bc) { bc = counts[c0 + c]; best = c; }
      cj_s = best;
    }
    __syncthreads();
    const int32_t cj = cj_s;
    for (uint32_t t = threadIdx.x; t < low_dim; t += blockDim.x) {
      const float sign = (t % 2 == 0) ? 1.0f + eps : 1.0f - eps;
      const float v = new_centroids[(c0 + cj) * low_dim + t];
      new_centroids[(c0 + ci) * low_dim + t] = v * sign;
      new_centroids[(c0 + cj) * low_dim + t] = v * (2.0f - sign);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double half = counts[c0 + cj] / 2.0;
      counts[c0 + ci] = half;
      counts[c0 + cj] -= half;
    }
    __syncthreads();
  }
  for (int32_t c = 0; c < L; ++c) {
    double ss = 0.0;
    for (uint32_t t = threadIdx.x; t < low_dim; t += blockDim.x) {
      const double v = static_cast<double>(new_centroids[(c0 + c) * low_dim + t]);
      ss += v * v;
    }
    red[threadIdx.x] = ss;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double nrm = red[0] > 0.0 ? sqrt(red[0]) : 1.0;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < low_dim; t += blockDim.x)
      centroids[(c0 + c) * low_dim + t] =
          static_cast<float>(static_cast<double>(new_centroids[(c0 + c) * low_dim + t]) / nrm);
  }
}

// Final assignment + probe list: float64 inner products, best-first insertion
// into a warp-resident list (lane j holds the j-th best so far).
__global__ void __launch_bounds__(256)
ivf_assign_kernel(const float* __restrict__ x, int64_t ld, int64_t n, uint32_t low_dim,
                  const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                  const int32_t* __restrict__ nlist, const int32_t* __restrict__ nprobe,
                  const int64_t* __restrict__ centroid_ptr, const float* __restrict__ centroids,
                  int32_t max_nprobe, const uint16_t* __restrict__ ell_idx, const float* __restrict__ ell_val,
                  int32_t W, int32_t* __restrict__ list_id, int32_t* __restrict__ probes) {
  extern __shared__ float smem_x[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (i >= n) return;
  const int64_t b = find_segment(bucket_ptr, n_buckets, i);
  const int32_t L = nlist[b];
  if (L == 0) {
    if (lane == 0) list_id[i] = 0;
    if (lane < max_nprobe) probes[i * max_nprobe + lane] = lane == 0 ? 0 : -1;
    return;
  }
  const int32_t P = nprobe[b];
  float* xi = smem_x + static_cast<size_t>(warp) * low_dim;
  if (ell_idx == nullptr) {
    for (uint32_t t = lane; t < low_dim; t += 32) xi[t] = x[i * ld + t];
    __syncwarp();
  }
  const float* cent = centroids + centroid_ptr[b] * low_dim;
  double my_score = -INFINITY;
  int32_t my_id = -1;
  for (int32_t c = 0; c < L; ++c) {
    const float* cr = cent + static_cast<int64_t>(c) * low_dim;
    double acc = 0.0;
    if (ell_idx != nullptr) {
      // sparse row: only the non-zero columns contribute (zero products are exact)
      for (int32_t j = lane; j < W; j += 32) {
        const float v = __ldg(ell_val + i * W + j);
        if (v != 0.f)
          acc = fma(static_cast<double>(v), static_cast<double>(__ldg(cr + __ldg(ell_idx + i * W + j))), acc);
      }
    } else {
      for (uint32_t t = lane; t < low_dim; t += 32)
        acc = fma(static_cast<double>(xi[t]), static_cast<double>(__ldg(cr + t)), acc);
    }
    acc = warp_sum_f64(acc);
    // entries ahead of the newcomer: strictly better, or equal (earlier id wins)
    const uint32_t ahead = __ballot_sync(0xffffffffu, my_id >= 0 && my_score >= acc);
    const int pos = __popc(ahead);
    if (pos < P) {
      const double up_s = __shfl_up_sync(0xffffffffu, my_score, 1);
      const int32_t up_i = __shfl_up_sync(0xffffffffu, my_id, 1);
      if (lane > pos) { my_score = up_s; my_id = up_i; }
      if (lane == pos) { my_score = acc; my_id = c; }
    }
  }
  if (lane < max_nprobe) probes[i * max_nprobe + lane] = lane < P ? my_id : -1;
  if (lane == 0) list_id[i] = my_id;
}


// ---------------------------------------------------------------- fused small-bucket trainer
// One CTA per bucket (persistent over buckets): the bucket's sparse rows stay in
// shared memory for all iterations, so HBM sees each row once.
//   assign: one THREAD per row -- sparse dot products against the centroids in
//     shared memory ([column][list] so one 16-byte load serves four lists),
//     float32, arg-max with ties to the lower id; no shuffles, no idle lanes;
//   update: 8 / L warps per list (each owning an interleaved share of the
//     columns) walk the list's rows in row order; a row's non-zero columns are
//     distinct, so lanes add them into the float64 accumulator without atomics
//     (shared-memory float atomics are CAS loops on sm_100) and in exactly the
//     oracle's summation order;
//   mean + L2 normalisation: one warp per list, no block-wide reductions;
//     empty lists (rare) take a slow path that re-seeds them from the largest
//     list with the +-1/1024 perturbation faiss uses;
//   after the last iteration the final assignment / probe list (float64 inner
//     products, ties to the lower id) is produced from the same shared memory.
// Launched per size class [need_lo, need_hi) so that small buckets run two CTAs per SM.
constexpr int kFusedMaxProbe = 4;

__global__ void __launch_bounds__(256)
kmeans_fused_kernel(const uint16_t* __restrict__ ell_idx, const float* __restrict__ ell_val, int32_t W,
                    uint32_t low_dim, const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                    const int32_t* __restrict__ nlist, const int64_t* __restrict__ centroid_ptr, int niter,
                    size_t need_lo, size_t need_hi, float* __restrict__ centroids,
                    const int32_t* __restrict__ nprobe, int32_t max_nprobe, int32_t* __restrict__ list_id,
                    int32_t* __restrict__ probes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int32_t cj_s;
  __shared__ double red_s[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = 8;
  const float eps = 1.0f / 1024.0f;
  const int d = static_cast<int>(low_dim);
  const int vp = W + 1;  // value pitch (odd: thread-per-row reads are conflict free)
  const int ip = W + 2;  // column pitch ((W + 2) / 2 is odd for W % 8 == 0)

  for (int64_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    const int32_t L = nlist[b];
    const int64_t s = bucket_ptr[b];
    const int32_t nb = static_cast<int32_t>(bucket_ptr[b + 1] - s);
    if (L <= 0) {
      // flat bucket: a single implicit list (written once, by the first size class)
      if (list_id != nullptr && need_lo == 0) {
        for (int t = tid; t < nb; t += 256) list_id[s + t] = 0;
        for (int t = tid; t < nb * max_nprobe; t += 256)
          probes[s * max_nprobe + t] = (t % max_nprobe) == 0 ? 0 : -1;
      }
      continue;
    }
    const size_t need = fused_smem_bytes(nb, L, W, low_dim);
    if (need < need_lo || need >= need_hi) continue;
    const int64_t c0 = centroid_ptr[b];
    const int lp = (L + 3) & ~3;
    // carve shared memory
    float* sval = reinterpret_cast<float*>(smem_raw);
    uint16_t* sidx = reinterpret_cast<uint16_t*>(smem_raw + static_cast<size_t>(nb) * vp * 4);
    size_t off = (static_cast<size_t>(nb) * vp * 4 + static_cast<size_t>(nb) * ip * 2 + 15) & ~size_t(15);
    double* acc = reinterpret_cast<double*>(smem_raw + off);
    off += static_cast<size_t>(L) * d * 8;
    float* C = reinterpret_cast<float*>(smem_raw + off);
    off += static_cast<size_t>(d) * lp * 4;
    double* cnt = reinterpret_cast<double*>(smem_raw + off);
    off += static_cast<size_t>(L) * 8;
    uint8_t* assign = smem_raw + off;
    __syncthreads();  // previous bucket fully written out
    for (int t = tid; t < nb * W; t += 256) {
      const int r = t / W, j = t - r * W;
      sval[r * vp + j] = __ldg(ell_val + s * W + t);
      sidx[r * ip + j] = __ldg(ell_idx + s * W + t);
    }
    for (int t = tid; t < d * lp; t += 256) C[t] = 0.f;
    __syncthreads();
    // init: centroid c = row (c * nb) / L
    for (int t = tid; t < L * W; t += 256) {
      const int c = t / W, j = t - c * W;
      const int row = static_cast<int>((static_cast<int64_t>(c) * nb) / L);
      const float v = sval[row * vp + j];
      if (v != 0.f) C[sidx[row * ip + j] * lp + c] = v;
    }
    __syncthreads();
    const int wpc = L >= kWarps ? 1 : kWarps / L;  // warps per list in the update (L is a power of two)
    for (int it = 0; it < niter; ++it) {
      // ---- assign (thread per row)
      for (int r = tid; r < nb; r += 256) {
        float best = -INFINITY;
        int best_c = 0;
        const float* rv = sval + r * vp;
        const uint16_t* ri = sidx + r * ip;
        for (int cb = 0; cb < lp; cb += 4) {
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
          for (int j = 0; j < W; ++j) {
            const float v = rv[j];
            const float4 cv = *reinterpret_cast<const float4*>(C + ri[j] * lp + cb);
            a0 = fmaf(v, cv.x, a0);
            a1 = fmaf(v, cv.y, a1);
            a2 = fmaf(v, cv.z, a2);
            a3 = fmaf(v, cv.w, a3);
          }
          if (cb + 0 < L && a0 > best) { best = a0; best_c = cb + 0; }
          if (cb + 1 < L && a1 > best) { best = a1; best_c = cb + 1; }
          if (cb + 2 < L && a2 > best) { best = a2; best_c = cb + 2; }
          if (cb + 3 < L && a3 > best) { best = a3; best_c = cb + 3; }
        }
        assign[r] = static_cast<uint8_t>(best_c);
      }
      for (int t = tid; t < L * d; t += 256) acc[t] = 0.0;
      __syncthreads();
      // ---- update: warp (c, h) adds the columns with idx % wpc == h of list c's rows, in row order
      for (int cw = warp; cw < L * wpc; cw += kWarps) {
        const int c = cw / wpc, h = cw - c * wpc;
        double* ac = acc + c * d;
        int n_c = 0;
        for (int r0 = 0; r0 < nb; r0 += 32) {
          const int rr = r0 + lane;
          uint32_t mine = __ballot_sync(0xffffffffu, rr < nb && assign[rr] == c);
          n_c += __popc(mine);
          while (mine) {
            const int r = r0 + __ffs(mine) - 1;
            mine &= mine - 1;
            for (int j = lane; j < W; j += 32) {
              const float v = sval[r * vp + j];
              const int k = sidx[r * ip + j];
              if (v != 0.f && (k & (wpc - 1)) == h) ac[k] += static_cast<double>(v);
            }
            __syncwarp();
          }
        }
        if (lane == 0 && h == 0) cnt[c] = static_cast<double>(n_c);
      }
      __syncthreads();
      bool any_empty = false;
      for (int c = 0; c < L; ++c) any_empty |= !(cnt[c] > 0.0);
      if (!any_empty) {
        // ---- fast path: mean + normalise, one warp per list
        for (int c = warp; c < L; c += kWarps) {
          const double inv_n = 1.0 / cnt[c];
          double ss = 0.0;
          for (int k = lane; k < d; k += 32) {
            const float m = static_cast<float>(acc[c * d + k] * inv_n);
            ss = fma(static_cast<double>(m), static_cast<double>(m), ss);
          }
          ss = warp_sum_f64(ss);
          const double inv_nrm = ss > 0.0 ? 1.0 / sqrt(ss) : 1.0;
          for (int k = lane; k < d; k += 32) {
            const float m = static_cast<float>(acc[c * d + k] * inv_n);
            C[k * lp + c] = static_cast<float>(static_cast<double>(m) * inv_nrm);
          }
        }
        __syncthreads();
      } else {
        // ---- slow path: mean (empty lists keep their centroid), re-seed, normalise
        for (int t = tid; t < L * d; t += 256) {
          const int c = t / d, k = t - c * d;
          const double n_c = cnt[c];
          if (n_c > 0.0) C[k * lp + c] = static_cast<float>(acc[t] * (1.0 / n_c));
        }
        __syncthreads();
        for (int ci = 0; ci < L; ++ci) {
          if (cnt[ci] > 0.0) continue;  // uniform
          if (tid == 0) {
            int bestc = 0;
            double bc = cnt[0];
            for (int c = 1; c < L; ++c)
              if (cnt[c] > bc) { bc = cnt[c]; bestc = c; }
            cj_s = bestc;
          }
          __syncthreads();
          const int cj = cj_s;
          for (int k = tid; k < d; k += 256) {
            const float sign = (k % 2 == 0) ? 1.0f + eps : 1.0f - eps;
            const float v = C[k * lp + cj];
            C[k * lp + ci] = v * sign;
            C[k * lp + cj] = v * (2.0f - sign);
          }
          __syncthreads();
          if (tid == 0) {
            const double half = cnt[cj] / 2.0;
            cnt[ci] = half;
            cnt[cj] -= half;
          }
          __syncthreads();
        }
        for (int c = 0; c < L; ++c) {
          double ss = 0.0;
          for (int k = tid; k < d; k += 256) {
            const double v = static_cast<double>(C[k * lp + c]);
            ss += v * v;
          }
          ss = warp_sum_f64(ss);
          if (lane == 0) red_s[warp] = ss;
          __syncthreads();
          double tot = 0.0;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) tot += red_s[w];
          const double inv_nrm = tot > 0.0 ? 1.0 / sqrt(tot) : 1.0;
          for (int k = tid; k < d; k += 256)
            C[k * lp + c] = static_cast<float>(static_cast<double>(C[k * lp + c]) * inv_nrm);
          __syncthreads();
        }
      }
    }
    for (int t = tid; t < L * d; t += 256) {
      const int c = t / d, k = t - c * d;
      centroids[c0 * low_dim + t] = C[k * lp + c];
    }
    // ---- final assignment + probe list (thread per row, float64)
    if (list_id != nullptr) {
      const int P = nprobe[b];
      for (int r = tid; r < nb; r += 256) {
        double ps[kFusedMaxProbe];
        int pi[kFusedMaxProbe];
#pragma unroll
        for (int t = 0; t < kFusedMaxProbe; ++t) { ps[t] = -INFINITY; pi[t] = -1; }
        const float* rv = sval + r * vp;
        const uint16_t* ri = sidx + r * ip;
        for (int cb = 0; cb < lp; cb += 4) {
          double a[4] = {0.0, 0.0, 0.0, 0.0};
          for (int j = 0; j < W; ++j) {
            const float v = rv[j];
            if (v != 0.f) {
              const float4 cv = *reinterpret_cast<const float4*>(C + ri[j] * lp + cb);
              const double dv = static_cast<double>(v);
              a[0] = fma(dv, static_cast<double>(cv.x), a[0]);
              a[1] = fma(dv, static_cast<double>(cv.y), a[1]);
              a[2] = fma(dv, static_cast<double>(cv.z), a[2]);
              a[3] = fma(dv, static_cast<double>(cv.w), a[3]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = cb + u;
            if (c < L) {
              // insert behind every entry that is better or equal (earlier id wins ties)
              const double sc = a[u];
              int pos = P;
#pragma unroll
              for (int t = 0; t < kFusedMaxProbe; ++t)
                if (t < P && pos == P && (pi[t] < 0 || sc > ps[t])) pos = t;
#pragma unroll
              for (int t = kFusedMaxProbe - 1; t >= 1; --t)
                if (t < P && t > pos) { ps[t] = ps[t - 1]; pi[t] = pi[t - 1]; }
#pragma unroll
              for (int t = 0; t < kFusedMaxProbe; ++t)
                if (t == pos) { ps[t] = sc; pi[t] = c; }
            }
          }
        }
        list_id[s + r] = pi[0];
        for (int t = 0; t < max_nprobe; ++t)
          probes[(s + r) * max_nprobe + t] = (t < P && t < kFusedMaxProbe) ? pi[t] : -1;
      }
    }
  }
}

struct KmeansLayout {
  int32_t* assign;
  float* new_centroids;
  double* counts;
};

static void kmeans_layout(Workspace& ws, int64_t n, int64_t total, uint32_t low_dim, KmeansLayout& L) {
  L.assign = ws.take<int32_t>(n);
  L.new_centroids = ws.take<float>(static_cast<size_t>(total) * low_dim);
  L.counts = ws.take<double>(total);
}

}  // namespace flc

extern "C" {

int flc_ivf_plan(const int64_t* bucket_ptr, int64_t n_buckets, int32_t n_probe, int exhaustive,
                 int32_t* nlist, int32_t* nprobe, int64_t* centroid_ptr, int64_t* total_centroids,
                 int32_t* max_nprobe, int64_t* max_ivf_bucket, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n_buckets >= 0, "n_buckets must be non-negative");
  FLC_REQUIRE(n_probe >= 1, "n_probe must be >= 1");
  FLC_REQUIRE(total_centroids && max_nprobe, "null host outputs");
  cudaStream_t stream = as_stream(stream_);
  timed("ivf_plan", stream, [&] { ivf_plan_kernel<<<1, 1024, 0, stream>>>(bucket_ptr, n_buckets, n_probe, exhaustive, nlist, nprobe,
                                          centroid_ptr); });
  FLC_LAUNCH_CHECK();
  int64_t tail[3] = {0, 0, 0};
  FLC_CUDA(cudaMemcpyAsync(tail, centroid_ptr + n_buckets, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  FLC_CUDA(cudaStreamSynchronize(stream));
  *total_centroids = tail[0];
  *max_nprobe = static_cast<int32_t>(tail[1] > 0 ? tail[1] : 1);
  if (max_ivf_bucket) *max_ivf_bucket = tail[2];
  return FLC_OK;
}

size_t flc_kmeans_workspace_bytes(int64_t n, int64_t total_centroids, uint32_t low_dim) {
  flc::Workspace ws(nullptr, 0);
  flc::KmeansLayout L;
  flc::kmeans_layout(ws, n > 0 ? n : 1, total_centroids > 0 ? total_centroids : 1, low_dim, L);
  return ws.used + 256;
}

int flc_kmeans_train(const float* x, int64_t ld, int64_t n, uint32_t low_dim, const int64_t* bucket_ptr,
                     int64_t n_buckets, const int32_t* nlist, const int64_t* centroid_ptr,
                     int64_t total_centroids, int64_t max_ivf_bucket, int niter,
                     const uint16_t* ell_idx, const float* ell_val, int32_t ell_width, float* centroids,
                     const int32_t* nprobe, int32_t max_nprobe, int32_t* list_id, int32_t* probes,
                     int32_t* assigned, void* workspace, size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && niter >= 0, "bad sizes");
  FLC_REQUIRE(low_dim > 0 && low_dim <= 8192, "low_dim must be in [1, 8192]");
  FLC_REQUIRE((ell_idx == nullptr) == (ell_val == nullptr), "ell_idx and ell_val go together");
  if (assigned) *assigned = 0;
  if (n == 0 || total_centroids == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
  // Buckets whose sparse rows + centroids fit in shared memory train in the fused
  // kernel; larger ones (and everything when no ELL copy is given) in the
  // generic multi-kernel path on the dense rows.
  const int32_t W = ell_idx ? ell_width : 0;
  size_t fused_limit = 0;
  bool any_generic = true;
  if (W > 0) {
    FLC_REQUIRE((W % 8) == 0, "ell_width must be a multiple of 8");
    fused_limit = kFusedSmemCap;
    const int64_t nb_max = max_ivf_bucket > 0 ? max_ivf_bucket : n;
    const size_t need_max = fused_smem_bytes(nb_max, nlist_rule(nb_max), W, low_dim);
    if (need_max <= kFusedSmemCap) any_generic = false;  // every IVF bucket fits
    // The fused trainer also emits the final assignment when it covers every bucket
    // and the probe lists are short.
    const bool emit = !any_generic && list_id != nullptr && probes != nullptr && nprobe != nullptr &&
                      max_nprobe >= 1 && max_nprobe <= kFusedMaxProbe;
    if (assigned) *assigned = emit ? 1 : 0;
    const size_t hi_bytes = need_max < kFusedSmemCap ? need_max + 16 : kFusedSmemCap + 16;
    // two size classes: [0, small) at two CTAs per SM, [small, cap] at one
    const size_t bounds[3] = {0, std::min(kFusedSmemSmall, hi_bytes), hi_bytes};
    for (int cls = 0; cls < 2; ++cls) {
      if (bounds[cls] >= bounds[cls + 1]) continue;
      const size_t smem = bounds[cls + 1];
      FLC_CUDA(cudaFuncSetAttribute(kmeans_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
      const int64_t cap = static_cast<int64_t>(kNumSMs) * (cls == 0 ? 2 : 1);
      const int64_t grid = n_buckets < cap ? n_buckets : cap;
      timed("kmeans_fused", stream, [&] { kmeans_fused_kernel<<<static_cast<unsigned>(grid), 256, smem, stream>>>(
          ell_idx, ell_val, W, low_dim, bucket_ptr, n_buckets, nlist, centroid_ptr, niter, bounds[cls],
          bounds[cls + 1], centroids, nprobe, max_nprobe, emit ? list_id : nullptr, emit ? probes : nullptr); });
      FLC_LAUNCH_CHECK();
    }
  }
  if (!any_generic) return FLC_OK;
  FLC_REQUIRE(x != nullptr, "dense rows are needed for buckets too large for the fused trainer");
  Workspace ws(workspace, workspace_bytes);
  KmeansLayout L;
  kmeans_layout(ws, n, total_centroids, low_dim, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "kmeans workspace too small: need %zu", ws.used);
  const size_t smem = static_cast<size_t>(8) * low_dim * sizeof(float);
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  const unsigned cblocks = static_cast<unsigned>(total_centroids);
  timed("kmeans_init", stream, [&] { kmeans_init_kernel<<<cblocks, 128, 0, stream>>>(x, ld, low_dim, bucket_ptr, n_buckets, nlist, centroid_ptr,
                                                  total_centroids, W, fused_limit, centroids); });
  FLC_LAUNCH_CHECK();
  for (int it = 0; it < niter; ++it) {
    timed("kmeans_assign", stream, [&] { kmeans_assign_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, stream>>>(
        x, ld, n, low_dim, bucket_ptr, n_buckets, nlist, centroid_ptr, centroids, W, fused_limit, L.assign); });
    FLC_LAUNCH_CHECK();
    timed("kmeans_update", stream, [&] { kmeans_update_kernel<<<cblocks, 128, 0, stream>>>(x, ld, low_dim, bucket_ptr, n_buckets, centroid_ptr,
                                                      total_centroids, L.assign, centroids, nlist, W, fused_limit,
                                                      L.new_centroids, L.counts); });
    FLC_LAUNCH_CHECK();
    timed("kmeans_fix", stream, [&] { kmeans_fix_kernel<<<static_cast<unsigned>(n_buckets), 128, 0, stream>>>(
        low_dim, n_buckets, nlist, centroid_ptr, bucket_ptr, W, fused_limit, L.new_centroids, L.counts,
        centroids); });
    FLC_LAUNCH_CHECK();
  }
  return FLC_OK;
}

int flc_ivf_assign(const float* x, int64_t ld, int64_t n, uint32_t low_dim, const int64_t* bucket_ptr,
                   int64_t n_buckets, const int32_t* nlist, const int32_t* nprobe,
                   const int64_t* centroid_ptr, const float* centroids, int32_t max_nprobe,
                   const uint16_t* ell_idx, const float* ell_val, int32_t ell_width,
                   int32_t* list_id, int32_t* probes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0, "bad n");
  FLC_REQUIRE(max_nprobe >= 1, "max_nprobe must be >= 1");
  FLC_REQUIRE((ell_idx == nullptr) == (ell_val == nullptr), "ell_idx and ell_val go together");
  FLC_REQUIRE(ell_idx != nullptr || x != nullptr, "need dense or ELL rows");
  if (max_nprobe > 32)
    return set_error(FLC_ERR_UNSUPPORTED, "n_probe > 32 is not supported by the device probe selection");
  if (n == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
  const size_t smem = static_cast<size_t>(8) * low_dim * sizeof(float);
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(ivf_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  timed("ivf_assign", stream, [&] { ivf_assign_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, stream>>>(
      x, ld, n, low_dim, bucket_ptr, n_buckets, nlist, nprobe, centroid_ptr, centroids, max_nprobe, ell_idx,
      ell_val, ell_width, list_id, probes); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
