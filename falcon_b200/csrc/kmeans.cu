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

// Single CTA: nlist / nprobe per bucket, exclusive scan -> centroid_ptr, max nprobe.
__global__ void __launch_bounds__(1024)
ivf_plan_kernel(const int64_t* __restrict__ bucket_ptr, int64_t n_buckets, int32_t n_probe, int exhaustive,
                int32_t* __restrict__ nlist, int32_t* __restrict__ nprobe, int64_t* __restrict__ centroid_ptr) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  __shared__ int32_t max_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { carry_s = 0; max_s = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n_buckets; base += 1024) {
    const int64_t b = base + tid;
    int32_t L = 0, P = 0;
    if (b < n_buckets) {
      L = nlist_rule(bucket_ptr[b + 1] - bucket_ptr[b]);
      if (L > 0) P = exhaustive ? L : max(1, min((L + 7) / 8, n_probe));
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
                                   int64_t total, float* __restrict__ centroids) {
  const int64_t gc = blockIdx.x;
  if (gc >= total) return;
  const int64_t b = find_segment(centroid_ptr, n_buckets, gc);
  const int64_t c = gc - centroid_ptr[b];
  const int64_t s = bucket_ptr[b], nb = bucket_ptr[b + 1] - s;
  const int64_t row = s + (c * nb) / nlist[b];
  for (uint32_t i = threadIdx.x; i < low_dim; i += blockDim.x)
    centroids[gc * low_dim + i] = x[row * ld + i];
}

__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ x, int64_t ld, int64_t n, uint32_t low_dim,
                     const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                     const int32_t* __restrict__ nlist, const int64_t* __restrict__ centroid_ptr,
                     const float* __restrict__ centroids, int32_t* __restrict__ assign) {
  extern __shared__ float smem_x[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (i >= n) return;
  const int64_t b = find_segment(bucket_ptr, n_buckets, i);
  const int32_t L = nlist[b];
  if (L == 0) {
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
                     float* __restrict__ new_centroids, double* __restrict__ counts) {
  const int64_t gc = blockIdx.x;
  if (gc >= total) return;
  const int64_t b = find_segment(centroid_ptr, n_buckets, gc);
  const int32_t c = static_cast<int32_t>(gc - centroid_ptr[b]);
  const int64_t s = bucket_ptr[b], e = bucket_ptr[b + 1];
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
                  const int64_t* __restrict__ centroid_ptr, float* __restrict__ new_centroids,
                  double* __restrict__ counts, float* __restrict__ centroids) {
  const int64_t b = blockIdx.x;
  if (b >= n_buckets) return;
  const int32_t L = nlist[b];
  if (L == 0) return;
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
                  int32_t max_nprobe, int32_t* __restrict__ list_id, int32_t* __restrict__ probes) {
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
  for (uint32_t t = lane; t < low_dim; t += 32) xi[t] = x[i * ld + t];
  __syncwarp();
  const float* cent = centroids + centroid_ptr[b] * low_dim;
  double my_score = -INFINITY;
  int32_t my_id = -1;
  for (int32_t c = 0; c < L; ++c) {
    const float* cr = cent + static_cast<int64_t>(c) * low_dim;
    double acc = 0.0;
    for (uint32_t t = lane; t < low_dim; t += 32)
      acc = fma(static_cast<double>(xi[t]), static_cast<double>(__ldg(cr + t)), acc);
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
                 int32_t* max_nprobe, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n_buckets >= 0, "n_buckets must be non-negative");
  FLC_REQUIRE(n_probe >= 1, "n_probe must be >= 1");
  FLC_REQUIRE(total_centroids && max_nprobe, "null host outputs");
  cudaStream_t stream = as_stream(stream_);
  timed("ivf_plan", stream, [&] { ivf_plan_kernel<<<1, 1024, 0, stream>>>(bucket_ptr, n_buckets, n_probe, exhaustive, nlist, nprobe,
                                          centroid_ptr); });
  FLC_LAUNCH_CHECK();
  int64_t tail[2] = {0, 0};
  FLC_CUDA(cudaMemcpyAsync(tail, centroid_ptr + n_buckets, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  FLC_CUDA(cudaStreamSynchronize(stream));
  *total_centroids = tail[0];
  *max_nprobe = static_cast<int32_t>(tail[1] > 0 ? tail[1] : 1);
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
                     int64_t total_centroids, int niter, float* centroids, void* workspace,
                     size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && niter >= 0, "bad sizes");
  FLC_REQUIRE(low_dim > 0 && low_dim <= 8192, "low_dim must be in [1, 8192]");
  if (n == 0 || total_centroids == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
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
                                                  total_centroids, centroids); });
  FLC_LAUNCH_CHECK();
  for (int it = 0; it < niter; ++it) {
    timed("kmeans_assign", stream, [&] { kmeans_assign_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, stream>>>(
        x, ld, n, low_dim, bucket_ptr, n_buckets, nlist, centroid_ptr, centroids, L.assign); });
    FLC_LAUNCH_CHECK();
    timed("kmeans_update", stream, [&] { kmeans_update_kernel<<<cblocks, 128, 0, stream>>>(x, ld, low_dim, bucket_ptr, n_buckets, centroid_ptr,
                                                      total_centroids, L.assign, centroids, L.new_centroids,
                                                      L.counts); });
    FLC_LAUNCH_CHECK();
    timed("kmeans_fix", stream, [&] { kmeans_fix_kernel<<<static_cast<unsigned>(n_buckets), 128, 0, stream>>>(
        low_dim, n_buckets, nlist, centroid_ptr, L.new_centroids, L.counts, centroids); });
    FLC_LAUNCH_CHECK();
  }
  return FLC_OK;
}

int flc_ivf_assign(const float* x, int64_t ld, int64_t n, uint32_t low_dim, const int64_t* bucket_ptr,
                   int64_t n_buckets, const int32_t* nlist, const int32_t* nprobe,
                   const int64_t* centroid_ptr, const float* centroids, int32_t max_nprobe,
                   int32_t* list_id, int32_t* probes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0, "bad n");
  FLC_REQUIRE(max_nprobe >= 1, "max_nprobe must be >= 1");
  if (max_nprobe > 32)
    return set_error(FLC_ERR_UNSUPPORTED, "n_probe > 32 is not supported by the device probe selection");
  if (n == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
  const size_t smem = static_cast<size_t>(8) * low_dim * sizeof(float);
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(ivf_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  timed("ivf_assign", stream, [&] { ivf_assign_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, stream>>>(
      x, ld, n, low_dim, bucket_ptr, n_buckets, nlist, nprobe, centroid_ptr, centroids, max_nprobe, list_id,
      probes); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
