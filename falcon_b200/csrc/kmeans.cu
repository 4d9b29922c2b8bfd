// Stage 2: per-bucket inverted-file index -- sizing, spherical k-means training,
// coarse assignment and probe selection (faiss IndexIVFFlat over an
// IndexFlatIP quantiser, METRIC_INNER_PRODUCT).
//
// Training works on the sparse (ELL) rows flc_vectorize emits (~35 of 400 columns
// are populated).  One definition of the arithmetic, two schedules:
//   * fused (kmeans_fused_kernel): one CTA per bucket, the bucket's rows resident
//     in shared memory for all iterations -- HBM sees every row once.  Covers
//     buckets of up to ~600 rows / 8 lists (all of them at 1 M spectra).
//   * tiled (kmeans_tiled_*): larger buckets; rows stream from HBM/L2 once per
//     iteration, list sums are accumulated with 64-bit integer atomics.
// Arithmetic (oracle/ivf.py:kmeans_train follows the same):
//   - assignment during training: float32 inner products, arg-max, ties to the
//     lower list id;
//   - list sums in 2^-40 fixed point (int64): sum_r rint(x[r][k] * 2^40).  Integer
//     addition is associative, so the result does not depend on the schedule or
//     on the order atomics land in -- both paths and every run give the same bits;
//   - mean = float32(double(sum) * (2^-40 / count)); empty lists keep their
//     centroid and are then re-seeded from the largest list with the +-1/1024
//     perturbation faiss uses; every centroid is L2-normalised (spherical k-means
//     for the IP metric): float32(double(c) * (1 / sqrt(sum c^2)));
//   - iterations stop early at a fixed point (assignments unchanged, no empty
//     list): further iterations would reproduce the same centroids bit for bit.
// The final assignment / probe selection uses float64 inner products so that it
// agrees with the oracle wherever there is no exact tie.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kmeans_tc.cuh"

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

constexpr float kFixScaleF = 1099511627776.0f;     // 2^40
constexpr double kFixInv = 1.0 / 1099511627776.0;  // 2^-40
constexpr int kFusedMaxL = 8;                      // lists per bucket the fused trainer handles
constexpr uint8_t kClsTiled = 2, kClsFlat = 3;

// Fused classes: 0 = two CTAs of 256 threads per SM, 1 = one CTA of 512 threads.
__host__ __device__ constexpr size_t fused_smem_limit(int cls) { return cls == 0 ? 110 * 1024 : 222 * 1024; }
__host__ __device__ constexpr int fused_threads(int cls) { return cls == 0 ? 256 : 512; }

// Shared-memory plan of one bucket in the fused trainer.  What only depends on
// (nb, L) comes first, so it can be placed before the row pitch is known.
struct FusedPlan {
  uint32_t rnnz, dlist, assign, C, acc, val, idx, total;
  int lp, copies;
};
__host__ __device__ inline FusedPlan fused_plan(int nb, int pitch, int L, int d, int warps) {
  FusedPlan p;
  p.lp = L <= 4 ? 4 : 8;  // centroid row length in shared memory
  p.copies = max(1, warps / (L <= 2 ? 2 : (L <= 4 ? 4 : 8)));  // warps (and accumulator copies) per list
  uint32_t o = 0;
  p.dlist = o;  o += 4u * nb * L;  // per list: rows that entered / left it this iteration
  p.rnnz = o;   o += 2u * nb;
  p.assign = o; o += nb;
  o = (o + 15u) & ~15u;
  p.C = o;      o += 4u * d * p.lp;
  p.acc = o;    o += 8u * d * L * p.copies;
  p.val = o;    o += 4u * nb * pitch;
  p.idx = o;    o += 2u * nb * pitch;
  p.total = (o + 15u) & ~15u;
  return p;
}

// Single CTA: nlist / nprobe per bucket, exclusive scan -> centroid_ptr, max nprobe, max IVF bucket.
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


// Best-P list of one row's (score, list) stream, fed in ascending list order by a whole warp; order:
// score descending, ties to the lower (= earlier) list.  Up to 32 probes lane j holds the j-th best in
// registers; beyond that (n_probe > 32, buckets of >= 512 lists) the list lives in the warp's slice of shared
// memory: a score that does not beat the current P-th best is rejected by one compare, the others are
// inserted by a warp-wide shift.
struct ProbeList {
  double my_score;
  int32_t my_id;
  double* s;
  int32_t* ids;
  int cnt, P;
  double worst;
  bool big;
  __device__ __forceinline__ void init(int P_, double* s_, int32_t* ids_) {
    P = P_;
    big = P_ > 32;
    s = s_;
    ids = ids_;
    cnt = 0;
    worst = -INFINITY;
    my_score = -INFINITY;
    my_id = -1;
  }
  // all lanes pass the same (sc, cid)
  __device__ __forceinline__ void push(double sc, int32_t cid, int lane) {
    if (!big) {
      const uint32_t ahead = __ballot_sync(0xffffffffu, my_id >= 0 && my_score >= sc);
      const int pos = __popc(ahead);
      if (pos < P) {
        const double up_s = __shfl_up_sync(0xffffffffu, my_score, 1);
        const int32_t up_i = __shfl_up_sync(0xffffffffu, my_id, 1);
        if (lane > pos) { my_score = up_s; my_id = up_i; }
        if (lane == pos) { my_score = sc; my_id = cid; }
      }
      return;
    }
    if (cnt == P && !(sc > worst)) return;  // equal scores: the earlier list stays
    int pos = 0;
    for (int j0 = 0; j0 < cnt; j0 += 32) {
      const int j = j0 + lane;
      const uint32_t ahead = __ballot_sync(0xffffffffu, j < cnt && s[j] >= sc);
      pos += __popc(ahead);
      if (ahead != 0xffffffffu) break;
    }
    const int ncnt = min(cnt + 1, P);
    for (int hi = ncnt - 1; hi > pos; hi -= 32) {  // shift [pos, ncnt - 1) up by one, from the top
      const int j = hi - lane;
      double ts = 0.0;
      int32_t ti = 0;
      if (j > pos) { ts = s[j - 1]; ti = ids[j - 1]; }
      __syncwarp();
      if (j > pos) { s[j] = ts; ids[j] = ti; }
      __syncwarp();
    }
    if (lane == 0) { s[pos] = sc; ids[pos] = cid; }
    __syncwarp();
    cnt = ncnt;
    if (cnt == P) worst = s[P - 1];
    __syncwarp();  // every lane has read the new P-th best before the next push may overwrite it
  }
  // probes[0 .. max_nprobe) of the row and its list
  __device__ __forceinline__ void store(int32_t* probes_row, int32_t max_nprobe, int32_t* list_id_row, int lane) {
    if (!big) {
      if (lane < max_nprobe) probes_row[lane] = lane < P ? my_id : -1;
      for (int t = 32 + lane; t < max_nprobe; t += 32) probes_row[t] = -1;
      if (lane == 0) *list_id_row = my_id;
    } else {
      for (int t = lane; t < max_nprobe; t += 32) probes_row[t] = t < cnt ? ids[t] : -1;
      if (lane == 0) *list_id_row = ids[0];
    }
  }
};
__host__ __device__ inline size_t probe_list_bytes(int32_t max_nprobe) {  // per warp
  return max_nprobe > 32 ? ((static_cast<size_t>(max_nprobe) * 12 + 15) & ~size_t(15)) : 0;
}

// Final assignment + probe list: one warp per row, float64 inner products,
// best-first insertion into a warp-resident list (lane j holds the j-th best so
// far).  With `bclass` only rows of tiled buckets are handled.
__global__ void __launch_bounds__(256)
ivf_assign_kernel(const float* __restrict__ x, int64_t ld, int64_t n, uint32_t low_dim,
                  const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                  const int32_t* __restrict__ nlist, const int32_t* __restrict__ nprobe,
                  const int64_t* __restrict__ centroid_ptr, const float* __restrict__ centroids,
                  int32_t max_nprobe, const uint16_t* __restrict__ ell_idx, const float* __restrict__ ell_val,
                  int32_t W, const uint8_t* __restrict__ bclass, int32_t* __restrict__ list_id,
                  int32_t* __restrict__ probes) {
  extern __shared__ float smem_x[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (i >= n) return;
  const int64_t b = find_segment(bucket_ptr, n_buckets, i);
  if (bclass != nullptr && bclass[b] != kClsTiled) return;
  const int32_t L = nlist[b];
  if (L == 0) {
    if (lane == 0) list_id[i] = 0;
    for (int t = lane; t < max_nprobe; t += 32) probes[i * max_nprobe + t] = t == 0 ? 0 : -1;
    return;
  }
  const int32_t P = min(nprobe[b], max_nprobe);  // max_nprobe may be a caller-side bound (checked by the caller)
  float* xi = smem_x + static_cast<size_t>(warp) * low_dim;
  unsigned char* pl_mem = reinterpret_cast<unsigned char*>(smem_x + static_cast<size_t>(8) * low_dim) +
                          warp * probe_list_bytes(max_nprobe);
  ProbeList top;
  top.init(P, reinterpret_cast<double*>(pl_mem), reinterpret_cast<int32_t*>(pl_mem + static_cast<size_t>(max_nprobe) * 8));
  if (ell_idx == nullptr) {
    for (uint32_t t = lane; t < low_dim; t += 32) xi[t] = x[i * ld + t];
    __syncwarp();
  }
  const float* cent = centroids + centroid_ptr[b] * low_dim;
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
    top.push(acc, c, lane);  // entries ahead of the newcomer: strictly better, or equal (earlier id wins)
  }
  top.store(probes + i * max_nprobe, max_nprobe, list_id + i, lane);
}

// ---------------------------------------------------------------- classification
// One warp per bucket: which trainer takes it (by the shared memory its rows
// need), appended to that class's work queue.  Flat buckets (no IVF) get their
// trivial list_id / probes here.
struct TrainQueues {
  int32_t* ctr;     // [4] pull counters
  int32_t* cnt;     // [4] queue lengths
  int32_t* queue;   // [3][n_buckets]
  uint8_t* bclass;  // [n_buckets]
};

__global__ void __launch_bounds__(256)
kmeans_classify_kernel(const uint16_t* __restrict__ ell_nnz, uint32_t low_dim,
                       const int64_t* __restrict__ bucket_ptr, int64_t n_buckets,
                       const int32_t* __restrict__ nlist, TrainQueues q, int force_tiled, int32_t max_nprobe,
                       int32_t* __restrict__ list_id, int32_t* __restrict__ probes,
                       const int64_t* __restrict__ centroid_ptr, int64_t total_centroids) {
  const int lane = threadIdx.x & 31;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (b >= n_buckets) return;
  const int32_t L = nlist[b];
  const int64_t s = bucket_ptr[b];
  const int64_t nb = bucket_ptr[b + 1] - s;
  // total_centroids may be a caller-side bound that turned out too small (sync-free callers check the true
  // total afterwards and redo the batch): buckets whose centroids would not fit are left untrained
  if (L <= 0 || centroid_ptr[b] + L > total_centroids) {
    if (lane == 0) q.bclass[b] = kClsFlat;
    if (list_id != nullptr) {
      for (int64_t t = lane; t < nb; t += 32) list_id[s + t] = 0;
      for (int64_t t = lane; t < nb * max_nprobe; t += 32)
        probes[s * max_nprobe + t] = (t % max_nprobe) == 0 ? 0 : -1;
    }
    return;
  }
  int wb = 1;
  for (int64_t t = lane; t < nb; t += 32) wb = max(wb, static_cast<int>(__ldg(ell_nnz + s + t)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wb = max(wb, __shfl_xor_sync(0xffffffffu, wb, o));
  int cls = kClsTiled;
  if (L <= kFusedMaxL && nb <= 32767 && !force_tiled) {
    for (int c = 1; c >= 0; --c)
      if (fused_plan(static_cast<int>(nb), wb | 1, L, static_cast<int>(low_dim), fused_threads(c) / 32).total <= fused_smem_limit(c)) cls = c;
  }
  if (lane == 0) {
    q.bclass[b] = static_cast<uint8_t>(cls);
    const int32_t pos = atomicAdd(q.cnt + cls, 1);
    q.queue[static_cast<int64_t>(cls) * n_buckets + pos] = static_cast<int32_t>(b);
  }
}

// ---------------------------------------------------------------- fused trainer
template <int LP, typename T>
__device__ __forceinline__ void row_scores(const float* __restrict__ vr, const uint16_t* __restrict__ ir, int m,
                                           const float* __restrict__ C, T (&a)[LP]) {
#pragma unroll
  for (int u = 0; u < LP; ++u) a[u] = T(0);
#pragma unroll 4
  for (int j = 0; j < m; ++j) {
    const T v = static_cast<T>(vr[j]);
    const float* cr = C + static_cast<int>(ir[j]) * LP;
#pragma unroll
    for (int u = 0; u < LP; u += 4) {
      const float4 c4 = *reinterpret_cast<const float4*>(cr + u);
      a[u] = fma(v, static_cast<T>(c4.x), a[u]);
      a[u + 1] = fma(v, static_cast<T>(c4.y), a[u + 1]);
      a[u + 2] = fma(v, static_cast<T>(c4.z), a[u + 2]);
      a[u + 3] = fma(v, static_cast<T>(c4.w), a[u + 3]);
    }
  }
}

struct FusedArgs {
  const uint16_t* ell_idx;
  const float* ell_val;
  const uint16_t* ell_nnz;
  int32_t W;
  uint32_t low_dim;
  const int64_t* bucket_ptr;
  int64_t n_buckets;
  const int32_t* nlist;
  const int64_t* centroid_ptr;
  int niter;
  TrainQueues q;
  float* centroids;
  const int32_t* nprobe;
  int32_t max_nprobe;
  int32_t* list_id;
  int32_t* probes;
  unsigned long long* timing;  // nullable: per-phase cycle counters (FLC_KMEANS_TIMING=1)
};

// Per-phase cycle counters of the fused trainer (debug aid, thread 0 of every CTA):
// 0 queue + row populations, 1 row load, 2 init, 3 assign, 4 update, 5 means, 6 normalise,
// 7 final assignment + write-out, 8 iterations run, 9 buckets.
__device__ unsigned long long g_kmeans_timing[16];

struct PhaseClock {
  unsigned long long* out;
  long long t, t0;
  __device__ PhaseClock(unsigned long long* o) : out(threadIdx.x == 0 ? o : nullptr), t(0), t0(0) {
    if (out) t = t0 = clock64();
  }
  // slowest bucket so far: [10] cycles, [11] rows << 32 | lists << 8 (best effort, racy)
  __device__ void finish(int nb, int L) {
    if (out) {
      const unsigned long long total = static_cast<unsigned long long>(clock64() - t0);
      if (atomicMax(out + 10, total) < total)
        out[11] = (static_cast<unsigned long long>(nb) << 32) | (static_cast<unsigned long long>(L) << 8);
    }
  }
  __device__ void mark(int phase) {
    if (out) {
      const long long now = clock64();
      atomicAdd(out + phase, static_cast<unsigned long long>(now - t));
      t = now;
    }
  }
  __device__ void count(int slot, unsigned long long v) {
    if (out) atomicAdd(out + slot, v);
  }
};

struct FusedStatic {  // static shared memory of the fused kernel
  int32_t cnt[8];  // rows per list
  int32_t dn[8];   // entries in each list's delta list
  double scale[8];
  double cntd[8];
  double red[16 * 8];
  int32_t wmax[16];
  int32_t pick;
  int32_t qi;
};

// Everything one bucket needs once its rows are in shared memory.  LP = centroid
// row length in shared memory (4 or 8 lists), NT = threads of the CTA.
template <int LP, int NT>
__device__ void fused_train_bucket(const FusedArgs& A, const FusedPlan& pl, unsigned char* smem, FusedStatic& S,
                                   int nb, int L, int pitch, int64_t s, int64_t c0, int P, PhaseClock& clk) {
  constexpr int kWarps = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = static_cast<int>(A.low_dim);
  const uint16_t* rnnz = reinterpret_cast<const uint16_t*>(smem + pl.rnnz);
  uint32_t* dlist = reinterpret_cast<uint32_t*>(smem + pl.dlist);
  uint8_t* assign = smem + pl.assign;
  float* C = reinterpret_cast<float*>(smem + pl.C);
  long long* acc = reinterpret_cast<long long*>(smem + pl.acc);
  const float* val = reinterpret_cast<const float*>(smem + pl.val);
  const uint16_t* idx = reinterpret_cast<const uint16_t*>(smem + pl.idx);
  const float eps = 1.0f / 1024.0f;

  bool converged = false;
  for (int it = 0; it < A.niter; ++it) {
    // ---- assign: one thread per row.  A row that moves from list a to list b is
    // queued as "+row" on b's delta list and "-row" on a's: the sums are integers,
    // so they can be maintained incrementally and the order of the queue is irrelevant.
    int changed = 0;
    for (int r0 = warp * 32; r0 < nb; r0 += NT) {
      const int r = r0 + lane;
      int old_c = -1, new_c = -1, m = 0;
      if (r < nb) {
        float a[LP];
        m = rnnz[r];
        row_scores<LP, float>(val + r * pitch, idx + r * pitch, m, C, a);
        float best = -INFINITY, second = -INFINITY;
        int best_c = 0;
#pragma unroll
        for (int u = 0; u < LP; ++u) {
          if (u < L) {
            if (a[u] > best) { second = best; best = a[u]; best_c = u; }
            else if (a[u] > second) { second = a[u]; }
          }
        }
        // bit 7: the runner-up is far enough behind (>> float32 rounding) that the float64 final
        // assignment cannot pick another list
        const int clear = (best - second > 1e-4f) ? 0x80 : 0;
        const int prev = assign[r];
        const int prev_c = prev == 0xff ? -1 : (prev & 0x7f);
        assign[r] = static_cast<uint8_t>(best_c | clear);
        if (prev_c != best_c) {
          old_c = prev_c;
          new_c = best_c;
          changed = 1;
        }
      }
      if (__any_sync(0xffffffffu, new_c >= 0)) {
        const uint32_t below = (1u << lane) - 1u;
#pragma unroll
        for (int u = 0; u < LP; ++u) {
          if (u < L) {
            const uint32_t enter = __ballot_sync(0xffffffffu, new_c == u);
            const uint32_t leave = __ballot_sync(0xffffffffu, old_c == u);
            if ((enter | leave) != 0u) {
              const int ne = __popc(enter);
              int base = 0;
              if (lane == 0) {
                base = atomicAdd(S.dn + u, ne + __popc(leave));
                atomicAdd(S.cnt + u, ne - __popc(leave));
              }
              base = __shfl_sync(0xffffffffu, base, 0);
              const uint32_t e = static_cast<uint32_t>(r) | (static_cast<uint32_t>(m) << 16);
              if (new_c == u) dlist[u * nb + base + __popc(enter & below)] = e;
              if (old_c == u) dlist[u * nb + base + ne + __popc(leave & below)] = e | 0x8000u;
            }
          }
        }
      }
    }
    changed = __syncthreads_or(changed);
    clk.mark(3);
    bool any_empty = false;
#pragma unroll
    for (int u = 0; u < LP; ++u) any_empty |= (u < L && S.cnt[u] == 0);
    if (it > 0 && !changed && !any_empty) {  // fixed point
      converged = true;
      break;
    }
    if (tid < L) S.scale[tid] = S.cnt[tid] > 0 ? kFixInv / static_cast<double>(S.cnt[tid]) : 0.0;
    // ---- update: the warps split every list's delta list (`copies` warps per list, each
    // with its own accumulator copy, added up in the means phase).  Lanes take the row's
    // columns (distinct within a row, so no conflicts inside a step): slots lane and lane + 32
    // in one step; the operands -- already converted to fixed point -- of the next row are in
    // flight during this row's read-modify-write.
    {
      const int copies = pl.copies;
      const int c = warp / copies, part = warp - c * copies;
      if (c < L) {
        long long* acc_c = acc + (part * L + c) * d;
        const uint32_t* dl = dlist + c * nb;
        const int n_c = S.dn[c];
        int k0 = 0, k1 = 0;
        long long q0 = 0, q1 = 0;
        uint32_t e = 0;
        auto fetch = [&](int i) {
          e = dl[i];
          const int m = static_cast<int>(e >> 16);
          const int o = static_cast<int>(e & 0x7fffu) * pitch;
          const bool neg = (e & 0x8000u) != 0u;
          q0 = 0; q1 = 0;
          if (lane < m) {
            k0 = idx[o + lane];
            const long long q = __float2ll_rn(val[o + lane] * kFixScaleF);
            q0 = neg ? -q : q;
          }
          if (lane + 32 < m) {
            k1 = idx[o + lane + 32];
            const long long q = __float2ll_rn(val[o + lane + 32] * kFixScaleF);
            q1 = neg ? -q : q;
          }
        };
        if (part < n_c) fetch(part);
        for (int i = part; i < n_c; i += copies) {
          const uint32_t ce = e;
          const int ck0 = k0, ck1 = k1;
          const long long cq0 = q0, cq1 = q1;
          if (i + copies < n_c) fetch(i + copies);
          const int m = static_cast<int>(ce >> 16);
          if (lane < m) acc_c[ck0] += cq0;
          if (lane + 32 < m) acc_c[ck1] += cq1;
          if (m > 64) {  // uniform; rows wider than two slots per lane (rare)
            const int o = static_cast<int>(ce & 0x7fffu) * pitch;
            const bool neg = (ce & 0x8000u) != 0u;
            for (int j = 64 + lane; j < m; j += 32) {
              const long long q = __float2ll_rn(val[o + j] * kFixScaleF);
              acc_c[idx[o + j]] += neg ? -q : q;
            }
          }
          __syncwarp();  // two rows may share a column: keep their read-modify-writes apart
        }
      }
    }
    __syncthreads();
    clk.mark(4);
    clk.count(8, 1);
    // ---- means (one thread per column), partial square norms
    {
      double ssq[LP];
#pragma unroll
      for (int u = 0; u < LP; ++u) ssq[u] = 0.0;
      for (int k = tid; k < d; k += NT) {
        float m[LP];
#pragma unroll
        for (int u = 0; u < LP; u += 4) {
          const float4 t = *reinterpret_cast<const float4*>(C + k * LP + u);
          m[u] = t.x; m[u + 1] = t.y; m[u + 2] = t.z; m[u + 3] = t.w;
        }
#pragma unroll
        for (int u = 0; u < LP; ++u) {
          if (u < L) {
            long long sum = 0;
            for (int cp = 0; cp < pl.copies; ++cp) sum += acc[(cp * L + u) * d + k];
            if (S.cnt[u] > 0) m[u] = static_cast<float>(__ll2double_rn(sum) * S.scale[u]);
            ssq[u] = fma(static_cast<double>(m[u]), static_cast<double>(m[u]), ssq[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < LP; u += 4)
          *reinterpret_cast<float4*>(C + k * LP + u) = make_float4(m[u], m[u + 1], m[u + 2], m[u + 3]);
      }
#pragma unroll
      for (int u = 0; u < LP; ++u) {
        const double t = warp_sum_f64(ssq[u]);
        if (lane == 0) S.red[warp * 8 + u] = t;
      }
    }
    __syncthreads();
    clk.mark(5);
    if (any_empty) {
      // ---- slow path: re-seed every empty list from the (currently) largest one
      if (tid < L) S.cntd[tid] = static_cast<double>(S.cnt[tid]);
      __syncthreads();
      for (int ci = 0; ci < L; ++ci) {
        if (S.cntd[ci] > 0.0) continue;  // uniform
        if (tid == 0) {
          int bestc = 0;
          double bc = S.cntd[0];
          for (int c = 1; c < L; ++c)
            if (S.cntd[c] > bc) { bc = S.cntd[c]; bestc = c; }
          S.pick = bestc;
        }
        __syncthreads();
        const int cj = S.pick;
        for (int k = tid; k < d; k += NT) {
          const float sign = (k % 2 == 0) ? 1.0f + eps : 1.0f - eps;
          const float v = C[k * LP + cj];
          C[k * LP + ci] = v * sign;
          C[k * LP + cj] = v * (2.0f - sign);
        }
        __syncthreads();
        if (tid == 0) {
          const double half = S.cntd[cj] / 2.0;
          S.cntd[ci] = half;
          S.cntd[cj] -= half;
        }
        __syncthreads();
      }
      // square norms again (the re-seeded lists changed)
      double ssq[LP];
#pragma unroll
      for (int u = 0; u < LP; ++u) ssq[u] = 0.0;
      for (int k = tid; k < d; k += NT) {
#pragma unroll
        for (int u = 0; u < LP; ++u) {
          const double v = static_cast<double>(C[k * LP + u]);
          ssq[u] = fma(v, v, ssq[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < LP; ++u) {
        const double t = warp_sum_f64(ssq[u]);
        if (lane == 0) S.red[warp * 8 + u] = t;
      }
      __syncthreads();
    }
    if (tid < 8) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += S.red[w * 8 + tid];
      S.scale[tid] = (tid < L && tot > 0.0) ? 1.0 / sqrt(tot) : 1.0;
      S.dn[tid] = 0;  // every warp is past its delta list
    }
    __syncthreads();
    for (int k = tid; k < d; k += NT) {
#pragma unroll
      for (int u = 0; u < LP; u += 4) {
        float4 t = *reinterpret_cast<const float4*>(C + k * LP + u);
        t.x = static_cast<float>(static_cast<double>(t.x) * S.scale[u]);
        t.y = static_cast<float>(static_cast<double>(t.y) * S.scale[u + 1]);
        t.z = static_cast<float>(static_cast<double>(t.z) * S.scale[u + 2]);
        t.w = static_cast<float>(static_cast<double>(t.w) * S.scale[u + 3]);
        *reinterpret_cast<float4*>(C + k * LP + u) = t;
      }
    }
    __syncthreads();
    clk.mark(6);
  }
  // ---- centroids out: [list][column]
  for (int t = tid; t < L * d; t += NT) {
    const int c = t / d, k = t - c * d;
    A.centroids[c0 * A.low_dim + t] = C[k * LP + c];
  }
  // ---- final assignment + probe list (one thread per row, float64, ties to the lower id)
  if (A.list_id != nullptr) {
    // At a fixed point the last training assignment was made against these very centroids: rows
    // whose float32 runner-up was clearly behind keep that list without the float64 pass.
    const bool shortcut = converged && P == 1;
    for (int r = tid; r < nb; r += NT) {
      if (shortcut && (assign[r] & 0x80)) {
        int32_t* pr = A.probes + (s + r) * A.max_nprobe;
        pr[0] = assign[r] & 0x7f;
        for (int t = 1; t < A.max_nprobe; ++t) pr[t] = -1;
        A.list_id[s + r] = assign[r] & 0x7f;
        continue;
      }
      double a[LP];
      row_scores<LP, double>(val + r * pitch, idx + r * pitch, rnnz[r], C, a);
      uint32_t used = 0;
      int32_t* pr = A.probes + (s + r) * A.max_nprobe;
      for (int t = 0; t < P; ++t) {
        double best = -INFINITY;
        int best_c = -1;
#pragma unroll
        for (int u = 0; u < LP; ++u)
          if (u < L && !((used >> u) & 1u) && (best_c < 0 || a[u] > best)) { best = a[u]; best_c = u; }
        used |= 1u << best_c;
        pr[t] = best_c;
        if (t == 0) A.list_id[s + r] = best_c;
      }
      for (int t = P; t < A.max_nprobe; ++t) pr[t] = -1;
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)  // class 0 must keep two CTAs per SM: <= 128 registers
kmeans_fused_kernel(FusedArgs A, int cls) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ FusedStatic S;
  constexpr int kWarps = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = static_cast<int>(A.low_dim);
  const int W = A.W;
  const int n_queued = A.q.cnt[cls];
  const int32_t* queue = A.q.queue + static_cast<int64_t>(cls) * A.n_buckets;

  for (;;) {
    __syncthreads();  // the previous bucket is done with shared memory
    if (tid == 0) S.qi = atomicAdd(A.q.ctr + cls, 1);
    __syncthreads();
    const int qi = S.qi;
    if (qi >= n_queued) break;
    PhaseClock clk(A.timing);
    const int64_t b = queue[qi];
    const int L = A.nlist[b];
    const int64_t s = A.bucket_ptr[b];
    const int nb = static_cast<int>(A.bucket_ptr[b + 1] - s);
    const int64_t c0 = A.centroid_ptr[b];
    // ---- row populations, widest row -> pitch
    FusedPlan pl = fused_plan(nb, 1, L, d, kWarps);
    uint16_t* rnnz = reinterpret_cast<uint16_t*>(smem + pl.rnnz);
    int wb = 1;
    for (int t = tid; t < nb; t += NT) {
      const int m = min(static_cast<int>(__ldg(A.ell_nnz + s + t)), W);
      rnnz[t] = static_cast<uint16_t>(m);
      wb = max(wb, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wb = max(wb, __shfl_xor_sync(0xffffffffu, wb, o));
    if (lane == 0) S.wmax[warp] = wb;
    if (tid < 8) { S.cnt[tid] = 0; S.dn[tid] = 0; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wb = max(wb, S.wmax[w]);
    clk.mark(0);
    const int pitch = wb | 1;  // odd: one thread per row reads without bank conflicts
    pl = fused_plan(nb, pitch, L, d, kWarps);
    float* val = reinterpret_cast<float*>(smem + pl.val);
    uint16_t* idx = reinterpret_cast<uint16_t*>(smem + pl.idx);
    float* C = reinterpret_cast<float*>(smem + pl.C);
    long long* acc = reinterpret_cast<long long*>(smem + pl.acc);
    uint8_t* assign = smem + pl.assign;
    // ---- rows: coalesced 16-byte loads of 8 slots (four chunks per thread in flight before
    // the first store), scattered into the odd-pitch layout
    {
      const int cpr = W >> 3;
      const int items = nb * cpr;
      constexpr int kBatch = 4;
      for (int i0 = tid; i0 < items; i0 += NT * kBatch) {
        uint4 ki[kBatch];
        float4 v0[kBatch], v1[kBatch];
        int rr[kBatch], jj[kBatch], mm[kBatch];
#pragma unroll
        for (int t = 0; t < kBatch; ++t) {
          const int i = i0 + t * NT;
          mm[t] = 0;
          if (i < items) {
            rr[t] = i / cpr;
            jj[t] = (i - rr[t] * cpr) << 3;
            mm[t] = rnnz[rr[t]];
            if (jj[t] < mm[t]) {
              const int64_t g = (s + rr[t]) * W + jj[t];
              ki[t] = __ldg(reinterpret_cast<const uint4*>(A.ell_idx + g));
              v0[t] = __ldg(reinterpret_cast<const float4*>(A.ell_val + g));
              v1[t] = __ldg(reinterpret_cast<const float4*>(A.ell_val + g + 4));
            } else {
              mm[t] = 0;
            }
          }
        }
#pragma unroll
        for (int t = 0; t < kBatch; ++t) {
          if (mm[t] > 0) {
            const float vv[8] = {v0[t].x, v0[t].y, v0[t].z, v0[t].w, v1[t].x, v1[t].y, v1[t].z, v1[t].w};
            const uint32_t kk[4] = {ki[t].x, ki[t].y, ki[t].z, ki[t].w};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (jj[t] + u < mm[t]) {
                val[rr[t] * pitch + jj[t] + u] = vv[u];
                idx[rr[t] * pitch + jj[t] + u] = static_cast<uint16_t>((kk[u >> 1] >> ((u & 1) * 16)) & 0xffffu);
              }
            }
          }
        }
      }
    }
    for (int t = tid; t < d * pl.lp; t += NT) C[t] = 0.f;
    for (int t = tid; t < d * L * pl.copies; t += NT) acc[t] = 0;
    for (int t = tid; t < nb; t += NT) assign[t] = 0xff;
    __syncthreads();
    clk.mark(1);
    // ---- initial centroids: evenly strided rows
    for (int t = tid; t < L * wb; t += NT) {
      const int c = t / wb, j = t - c * wb;
      const int row = static_cast<int>((static_cast<int64_t>(c) * nb) / L);
      if (j < rnnz[row]) C[idx[row * pitch + j] * pl.lp + c] = val[row * pitch + j];
    }
    __syncthreads();
    const int P = A.list_id != nullptr ? min(min(A.nprobe[b], L), A.max_nprobe) : 0;
    clk.mark(2);
    if (pl.lp == 4)
      fused_train_bucket<4, NT>(A, pl, smem, S, nb, L, pitch, s, c0, P, clk);
    else
      fused_train_bucket<8, NT>(A, pl, smem, S, nb, L, pitch, s, c0, P, clk);
    __syncthreads();
    clk.mark(7);
    clk.count(9, 1);
    clk.finish(nb, L);
  }
}

// ---------------------------------------------------------------- tiled trainer (large buckets)
// Rows stream from global memory once per iteration; the bucket's centroids are
// staged in shared memory 32 lists at a time ([column][32], from the transposed
// training copy `ct`); sums go to global int64 accumulators with atomics.
// 16-byte shared-memory load by 32-bit shared address.  Used for the [column][32 lists] centroid
// tiles of the thread-per-row kernels: thread t reads the eight 16-byte chunks of a 128-byte row
// in the order chunk ^ (t & 7), so the eight threads of a quarter warp hit eight different bank
// groups whatever their columns are (a plain chunk order would be an eight-way conflict).
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

struct TiledArgs {
  const uint16_t* ell_idx;
  const float* ell_val;
  const uint16_t* ell_nnz;
  int32_t W;
  uint32_t low_dim;
  int64_t n;
  const int64_t* bucket_ptr;
  int64_t n_buckets;
  const int32_t* nlist;
  const int64_t* centroid_ptr;
  TrainQueues q;
  float* centroids;  // [total][d] final layout
  float* ct;         // per bucket [d][L]: training copy
  long long* gsum;   // [total][d]
  int32_t* gcnt;     // [total]
  double* gcntd;     // [total] (re-seeding of empty lists)
  int32_t* gassign;  // [n] previous assignment
  int32_t* bstate;   // [n_buckets][2]: changed flag, converged flag
  // tensor-core assignment (kmeans_tc.cu); cb == nullptr: SIMT assignment
  uint16_t* cb;      // [total][ld_c] bf16 copy of the centroids (TMA operand)
  int64_t ld_c;
  int32_t* tc_best;  // [n] arg-max list by bf16 scores; top bit: too close to call, re-scored by the apply kernel
  int32_t* tc_unsure;  // [n] rows the float32 pass of the final assignment could not decide (first *tc_counts[1])
  int4* units;       // query tiles of the tiled buckets still training
  int32_t* unit_bucket;
  int32_t* tc_counts;  // [0] units, [1] rows of the float64 final pass, [2] units of the sparse-row kernel, [3] of the dense-row kernel
  int4* units_sp;    // the same tiles, split by which tensor-core kernel scores them
  int4* units_dn;
};

__global__ void __launch_bounds__(128)
kmeans_tiled_init_kernel(TiledArgs A, int64_t total) {
  const int64_t gc = blockIdx.x;
  if (gc >= total) return;
  const int64_t b = find_segment(A.centroid_ptr, A.n_buckets, gc);
  if (A.q.bclass[b] != kClsTiled) return;
  const int64_t c0 = A.centroid_ptr[b];
  const int64_t c = gc - c0;
  const int32_t L = A.nlist[b];
  if (c >= L) return;  // `total` may be an upper bound of the centroid count (sync-free callers)
  const int64_t s = A.bucket_ptr[b], nb = A.bucket_ptr[b + 1] - s;
  const int64_t row = s + (c * nb) / L;
  const uint32_t d = A.low_dim;
  float* cr = A.centroids + gc * d;
  float* ctb = A.ct + c0 * d;
  uint16_t* cbr = A.cb ? A.cb + gc * A.ld_c : nullptr;
  for (uint32_t k = threadIdx.x; k < d; k += blockDim.x) {
    cr[k] = 0.f;
    ctb[static_cast<int64_t>(k) * L + c] = 0.f;
  }
  if (cbr)
    for (int64_t k = threadIdx.x; k < A.ld_c; k += blockDim.x) cbr[k] = 0;
  __syncthreads();
  const int m = min(static_cast<int>(A.ell_nnz[row]), A.W);
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    const uint32_t k = A.ell_idx[row * A.W + j];
    const float v = A.ell_val[row * A.W + j];
    cr[k] = v;
    ctb[static_cast<int64_t>(k) * L + c] = v;
    if (cbr) cbr[k] = f32_to_bf16_rne(v);
  }
}

// Unit descriptors of the tensor-core assignment: every tiled bucket contributes
// ceil(rows / 128) query tiles (first row, bucket end, first / end centroid row).
__global__ void kmeans_tc_units_kernel(TiledArgs A, int32_t min_lists, int32_t sparse_max_lists) {
  // one warp per queued bucket: lane 0 reserves the ranges, the lanes write the descriptors
  const int lane = threadIdx.x & 31;
  const int32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (qi >= A.q.cnt[kClsTiled]) return;
  const int64_t b = A.q.queue[static_cast<int64_t>(kClsTiled) * A.n_buckets + qi];
  if (A.bstate[2 * b + 1] != 0) return;  // converged: nothing left to do
  if (A.nlist[b] < min_lists) return;     // the thread-per-row kernel owns the buckets with few lists
  const int64_t s = A.bucket_ptr[b], e = A.bucket_ptr[b + 1];
  const int64_t c0 = A.centroid_ptr[b];
  const int32_t tq = static_cast<int32_t>((e - s + 127) / 128);
  const bool sparse = A.nlist[b] <= sparse_max_lists;
  int4* mine = sparse ? A.units_sp : A.units_dn;
  int32_t base = 0, mbase = 0;
  if (lane == 0) {
    base = atomicAdd(A.tc_counts, tq);
    mbase = atomicAdd(A.tc_counts + (sparse ? 2 : 3), tq);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  mbase = __shfl_sync(0xffffffffu, mbase, 0);
  for (int32_t t = lane; t < tq; t += 32) {
    const int4 ud = make_int4(static_cast<int>(s + 128 * t), static_cast<int>(e), static_cast<int>(c0),
                              static_cast<int>(c0 + A.nlist[b]));
    A.units[base + t] = ud;
    A.unit_bucket[base + t] = static_cast<int32_t>(b);
    mine[mbase + t] = ud;  // consecutive tiles of a bucket stay consecutive: the sparse kernel keeps the centroids resident
  }
}

// After the tensor-core pass: one CTA of 128 threads per query tile of the buckets still training.
// (1) Rows the bf16 scores could not decide (top bit of tc_best set) get the exact float32 arg-max
//     -- the fused trainer's arithmetic: products added in slot order with fmaf, ties to the lower
//     list -- warp-cooperatively, one lane per list; the tile's descriptor names the bucket's
//     centroids, so there is no search.
// (2) Rows whose list changed move their fixed-point values between the lists' sums: the warp
//     walks its changed rows and its lanes cover a row's slots (int64 global atomics).
__global__ void __launch_bounds__(128)
kmeans_tiled_apply_kernel(TiledArgs A) {
  const int lane = threadIdx.x & 31;
  const int d = static_cast<int>(A.low_dim);
  const int W = A.W;
  for (int64_t u = blockIdx.x; u < A.tc_counts[0]; u += gridDim.x) {
    const int4 ud = A.units[u];
    const int64_t warp_row0 = static_cast<int64_t>(ud.x) + (threadIdx.x & ~31);
    const int64_t i = warp_row0 + lane;
    const bool have = i < ud.y;
    const int64_t c0 = ud.z;
    const int32_t L = ud.w - ud.z;
    const float* ctb = A.ct + c0 * d;
    int new_c = have ? A.tc_best[i] : 0;
    uint32_t todo = __ballot_sync(0xffffffffu, have && new_c < 0);
    while (todo != 0u) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1u;
      const int64_t r = warp_row0 + src;
      const int m = min(static_cast<int>(A.ell_nnz[r]), W);
      float best = -INFINITY;
      int best_c = 0x7fffffff;
      for (int32_t s0 = 0; s0 < L; s0 += 32) {
        const int32_t c = s0 + lane;
        float acc = 0.f;
        for (int j0 = 0; j0 < m; j0 += 32) {
          const int j = j0 + lane;
          const uint32_t kj = j < m ? static_cast<uint32_t>(__ldg(A.ell_idx + r * W + j)) : 0u;
          const float vj = j < m ? __ldg(A.ell_val + r * W + j) : 0.f;
          const int cnt = min(32, m - j0);
          const int cl = min(c, L - 1);  // lanes past the last list read a valid address and are ignored below
          for (int t0 = 0; t0 < cnt; t0 += 8) {
            // eight independent gathers in flight, then the products in slot order
            float cv[8], vv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t k = __shfl_sync(0xffffffffu, kj, (t0 + e) & 31);
              vv[e] = __shfl_sync(0xffffffffu, vj, (t0 + e) & 31);
              cv[e] = __ldg(ctb + static_cast<int64_t>(k) * L + cl);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (t0 + e < cnt) acc = fmaf(vv[e], cv[e], acc);
          }
        }
        if (c < L && acc > best) { best = acc; best_c = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        if (ov > best || (ov == best && oc < best_c)) { best = ov; best_c = oc; }
      }
      if (lane == src) new_c = best_c;
    }
    const int old_c = have ? A.gassign[i] : 0;
    const bool changed = have && old_c != new_c;
    if (changed) {
      A.gassign[i] = new_c;
      atomicAdd(A.gcnt + c0 + new_c, 1);
      if (old_c >= 0) atomicSub(A.gcnt + c0 + old_c, 1);
    }
    uint32_t moved = __ballot_sync(0xffffffffu, changed);
    if (moved != 0u && lane == 0) A.bstate[2 * static_cast<int64_t>(A.unit_bucket[u])] = 1;  // benign race: every writer stores 1
    while (moved != 0u) {
      const int src = __ffs(moved) - 1;
      moved &= moved - 1u;
      const int nc = __shfl_sync(0xffffffffu, new_c, src);
      const int oc = __shfl_sync(0xffffffffu, old_c, src);
      const int64_t r = warp_row0 + src;
      const int m = min(static_cast<int>(A.ell_nnz[r]), W);
      unsigned long long* add = reinterpret_cast<unsigned long long*>(A.gsum + (c0 + nc) * d);
      unsigned long long* sub = reinterpret_cast<unsigned long long*>(A.gsum + (c0 + max(oc, 0)) * d);
      for (int j = lane; j < m; j += 32) {
        const uint32_t k = __ldg(A.ell_idx + r * W + j);
        const long long q = __float2ll_rn(__ldg(A.ell_val + r * W + j) * kFixScaleF);
        atomicAdd(add + k, static_cast<unsigned long long>(q));
        if (oc >= 0) atomicAdd(sub + k, static_cast<unsigned long long>(-q));
      }
    }
  }
}

constexpr int kTiledThreads = 256;
constexpr int kTiledRounds = 4;                              // rows per thread
constexpr int kTiledRows = kTiledThreads * kTiledRounds;     // rows per CTA
constexpr int kTiledG = 32;                                  // lists staged per pass

// Assignment + incremental accumulation.  A CTA owns 1024 consecutive rows; for
// every tiled bucket segment inside them it stages the centroids 32 lists at a
// time ([column][32], chunk-swizzled reads: see lds128) and each thread scores its
// rows from 16-byte row chunks: float32, products added in slot order with fmaf, ties
// to the lower list -- the fused trainer's arithmetic.
// Rows whose list changed move their fixed-point values between the lists' global
// int64 sums with atomics (integer: order free, exact).  Buckets with more than
// `max_lists` lists are left to the tensor-core assignment.
__global__ void __launch_bounds__(kTiledThreads, 2)
kmeans_tiled_assign_kernel(TiledArgs A, int32_t max_lists) {
  extern __shared__ __align__(128) float Cs[];  // [d][kTiledG]
  const int tid = threadIdx.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kTiledRows;
  const int64_t i1 = min(i0 + kTiledRows, A.n);
  const int d = static_cast<int>(A.low_dim);
  const int W = A.W;
  const uint32_t cs_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Cs));  // 128-byte aligned
  const uint32_t rot = static_cast<uint32_t>(tid & 7);
  int64_t b = find_segment(A.bucket_ptr, A.n_buckets, i0);
  for (; b < A.n_buckets && A.bucket_ptr[b] < i1; ++b) {
    if (A.q.bclass[b] != kClsTiled || A.bstate[2 * b + 1] != 0) continue;  // uniform
    const int32_t L = A.nlist[b];
    if (L > max_lists) continue;
    const int64_t s = max(A.bucket_ptr[b], i0), e = min(A.bucket_ptr[b + 1], i1);
    const int64_t c0 = A.centroid_ptr[b];
    const float* ctb = A.ct + c0 * d;
    float best[kTiledRounds];
    int best_c[kTiledRounds];
#pragma unroll
    for (int t = 0; t < kTiledRounds; ++t) { best[t] = -INFINITY; best_c[t] = 0x7fffffff; }
    for (int g0 = 0; g0 < L; g0 += kTiledG) {
      const int G = min(kTiledG, L - g0);
      __syncthreads();
      if (L == kTiledG && (reinterpret_cast<uintptr_t>(ctb) & 15) == 0) {
        const float4* src = reinterpret_cast<const float4*>(ctb);
        float4* dst = reinterpret_cast<float4*>(Cs);
        for (int t = tid; t < d * (kTiledG / 4); t += kTiledThreads) dst[t] = __ldg(src + t);
      } else {
        for (int t = tid; t < d * kTiledG; t += kTiledThreads) {
          const int k = t / kTiledG, g = t - k * kTiledG;
          Cs[t] = g < G ? __ldg(ctb + static_cast<int64_t>(k) * L + g0 + g) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int rnd = 0; rnd < kTiledRounds; ++rnd) {
        const int64_t i = s + rnd * kTiledThreads + tid;
        if (i >= e) continue;
        const int m = min(static_cast<int>(A.ell_nnz[i]), W);
        float a[kTiledG];
#pragma unroll
        for (int u = 0; u < kTiledG; ++u) a[u] = 0.f;
        for (int j0 = 0; j0 < m; j0 += 8) {
          const uint4 ki = __ldg(reinterpret_cast<const uint4*>(A.ell_idx + i * W + j0));
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(A.ell_val + i * W + j0));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(A.ell_val + i * W + j0 + 4));
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          const uint32_t kk[4] = {ki.x, ki.y, ki.z, ki.w};
#pragma unroll
          for (int t = 0; t < 8; ++t) {  // zero padding multiplies to zero
            const float v = vv[t];
            const uint32_t row = cs_addr + (((kk[t >> 1] >> ((t & 1) * 16)) & 0xffffu) * (kTiledG * 4u)) + rot * 16u;
#pragma unroll
            for (int u = 0; u < kTiledG; u += 4) {
              const float4 c4 = lds128(row ^ (u * 4u));  // chunk (u / 4) ^ rot of the row
              a[u] = fmaf(v, c4.x, a[u]);
              a[u + 1] = fmaf(v, c4.y, a[u + 1]);
              a[u + 2] = fmaf(v, c4.z, a[u + 2]);
              a[u + 3] = fmaf(v, c4.w, a[u + 3]);
            }
          }
        }
        // a[u + e] holds list g0 + (u ^ 4 rot) + e
#pragma unroll
        for (int u = 0; u < kTiledG; u += 4) {
          const int g = u ^ static_cast<int>(rot * 4u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = g0 + g + q;
            if (g + q < G && (a[u + q] > best[rnd] || (a[u + q] == best[rnd] && c < best_c[rnd]))) {
              best[rnd] = a[u + q];
              best_c[rnd] = c;
            }
          }
        }
      }
    }
#pragma unroll
    for (int rnd = 0; rnd < kTiledRounds; ++rnd) {
      const int64_t i = s + rnd * kTiledThreads + tid;
      if (i >= e) continue;
      const int old_c = A.gassign[i];
      const int new_c = best_c[rnd] == 0x7fffffff ? 0 : best_c[rnd];
      if (old_c == new_c) continue;
      A.gassign[i] = new_c;
      A.bstate[2 * b] = 1;  // benign race: every writer stores 1
      atomicAdd(A.gcnt + c0 + new_c, 1);
      if (old_c >= 0) atomicSub(A.gcnt + c0 + old_c, 1);
      unsigned long long* add = reinterpret_cast<unsigned long long*>(A.gsum + (c0 + new_c) * d);
      unsigned long long* sub = reinterpret_cast<unsigned long long*>(A.gsum + (c0 + max(old_c, 0)) * d);
      const int m = min(static_cast<int>(A.ell_nnz[i]), W);
      for (int j = 0; j < m; ++j) {
        const uint32_t k = __ldg(A.ell_idx + i * W + j);
        const long long q = __float2ll_rn(__ldg(A.ell_val + i * W + j) * kFixScaleF);
        atomicAdd(add + k, static_cast<unsigned long long>(q));
        if (old_c >= 0) atomicAdd(sub + k, static_cast<unsigned long long>(-q));
      }
    }
  }
}

// One CTA per tiled bucket: means, re-seeding of empty lists, normalisation; writes
// both centroid layouts, records convergence.
// Common case (no empty list): one WARP per list, eight lists in flight.  The mean goes to a
// [32 lists][d + 1] shared-memory tile while its sum of squares is accumulated, the second pass
// scales the tile in place and writes the row-major float32 and bf16 copies, and the transposed
// training copy ct[column][list] is then written straight from the tile (row stride d + 1 is odd:
// the column reads are conflict free, the global writes are coalesced).  Same arithmetic, in the
// same order, as the general path below, which stays for buckets with an empty list.
__global__ void __launch_bounds__(256)
kmeans_tiled_update_kernel(TiledArgs A) {
  extern __shared__ float mean_tile[];  // [32][d + 1]
  __shared__ int32_t cj_s;
  __shared__ int32_t empty_s;
  __shared__ float tile[32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = static_cast<int>(A.low_dim);
  const int ldt = d + 1;
  const float eps = 1.0f / 1024.0f;
  const int32_t n_queued = A.q.cnt[kClsTiled];
  const int32_t* queue = A.q.queue + static_cast<int64_t>(kClsTiled) * A.n_buckets;
  for (int32_t qi = blockIdx.x; qi < n_queued; qi += gridDim.x) {
    const int64_t b = queue[qi];
    if (A.bstate[2 * b + 1] != 0) continue;  // converged earlier
    const int32_t L = A.nlist[b];
    const int64_t c0 = A.centroid_ptr[b];
    float* cent = A.centroids + c0 * d;
    float* ctb = A.ct + c0 * d;
    long long* gs = A.gsum + c0 * d;
    if (tid == 0) empty_s = 0;
    __syncthreads();
    for (int32_t c = tid; c < L; c += 256)
      if (A.gcnt[c0 + c] <= 0) empty_s = 1;  // benign race: every writer stores 1
    __syncthreads();
    if (empty_s == 0) {
      for (int32_t ci = 0; ci < L; ci += 32) {
        const int32_t c_end = min(L, ci + 32);
        for (int32_t c = ci + warp; c < c_end; c += 8) {
          const double scale = kFixInv / static_cast<double>(A.gcnt[c0 + c]);
          const long long* gr = gs + static_cast<int64_t>(c) * d;
          float* tr = mean_tile + (c - ci) * ldt;
          double ss = 0.0;
#pragma unroll 4
          for (int k = lane; k < d; k += 32) {
            const float v = static_cast<float>(__ll2double_rn(gr[k]) * scale);
            tr[k] = v;
            const double vd = static_cast<double>(v);
            ss = fma(vd, vd, ss);
          }
          ss = warp_sum_f64(ss);
          const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 1.0;
          float* cr = cent + static_cast<int64_t>(c) * d;
          uint16_t* cbr = A.cb ? A.cb + (c0 + c) * A.ld_c : nullptr;
#pragma unroll 4
          for (int k = lane; k < d; k += 32) {
            const float v = static_cast<float>(static_cast<double>(tr[k]) * inv);
            tr[k] = v;
            cr[k] = v;
            if (cbr) cbr[k] = f32_to_bf16_rne(v);
          }
        }
        __syncthreads();
        const int32_t c = ci + lane;
        if (c < L)
          for (int k = warp; k < d; k += 8) ctb[static_cast<int64_t>(k) * L + c] = mean_tile[lane * ldt + k];
        __syncthreads();
      }
      if (tid == 0) {
        if (A.bstate[2 * b] == 0) A.bstate[2 * b + 1] = 1;
        A.bstate[2 * b] = 0;
      }
      __syncthreads();
      continue;
    }
    bool any_empty = false;
    for (int32_t c = 0; c < L; ++c) {
      const int32_t n_c = A.gcnt[c0 + c];
      if (n_c > 0) {
        const double scale = kFixInv / static_cast<double>(n_c);
        for (int k = tid; k < d; k += 256) {
          cent[static_cast<int64_t>(c) * d + k] =
              static_cast<float>(__ll2double_rn(gs[static_cast<int64_t>(c) * d + k]) * scale);
        }
      } else {
        any_empty = true;
      }
    }
    __syncthreads();
    if (any_empty) {
      for (int32_t c = tid; c < L; c += 256) A.gcntd[c0 + c] = static_cast<double>(A.gcnt[c0 + c]);
      __syncthreads();
      for (int32_t ci = 0; ci < L; ++ci) {
        if (A.gcntd[c0 + ci] > 0.0) continue;  // uniform
        if (tid == 0) {
          int32_t best = 0;
          double bc = A.gcntd[c0];
          for (int32_t c = 1; c < L; ++c)
            if (A.gcntd[c0 + c] > bc) { bc = A.gcntd[c0 + c]; best = c; }
          cj_s = best;
        }
        __syncthreads();
        const int32_t cj = cj_s;
        for (int k = tid; k < d; k += 256) {
          const float sign = (k % 2 == 0) ? 1.0f + eps : 1.0f - eps;
          const float v = cent[static_cast<int64_t>(cj) * d + k];
          cent[static_cast<int64_t>(ci) * d + k] = v * sign;
          cent[static_cast<int64_t>(cj) * d + k] = v * (2.0f - sign);
        }
        __syncthreads();
        if (tid == 0) {
          const double half = A.gcntd[c0 + cj] / 2.0;
          A.gcntd[c0 + ci] = half;
          A.gcntd[c0 + cj] -= half;
        }
        __syncthreads();
      }
    }
    // normalise: one warp per list (row-major float32 + bf16 copies, coalesced)
    for (int32_t c = warp; c < L; c += 8) {
      float* cr = cent + static_cast<int64_t>(c) * d;
      double ss = 0.0;
      for (int k = lane; k < d; k += 32) {
        const double v = static_cast<double>(cr[k]);
        ss = fma(v, v, ss);
      }
      ss = warp_sum_f64(ss);
      const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 1.0;
      uint16_t* cbr = A.cb ? A.cb + (c0 + c) * A.ld_c : nullptr;
      for (int k = lane; k < d; k += 32) {
        const float v = static_cast<float>(static_cast<double>(cr[k]) * inv);
        cr[k] = v;
        if (cbr) cbr[k] = f32_to_bf16_rne(v);
      }
    }
    __syncthreads();
    // transposed training copy ct[column][list] through 32 x 32 shared-memory tiles, so that
    // both the reads (rows of `cent`) and the writes (rows of `ct`) are coalesced
    for (int32_t ci = 0; ci < L; ci += 32) {
      for (int ki = 0; ki < d; ki += 32) {
        for (int r = warp; r < 32; r += 8) {
          const int32_t c = ci + r;
          const int k = ki + lane;
          tile[r][lane] = (c < L && k < d) ? cent[static_cast<int64_t>(c) * d + k] : 0.f;
        }
        __syncthreads();
        for (int r = warp; r < 32; r += 8) {
          const int k = ki + r;
          const int32_t c = ci + lane;
          if (k < d && c < L) ctb[static_cast<int64_t>(k) * L + c] = tile[lane][r];
        }
        __syncthreads();
      }
    }
    if (tid == 0) {  // sums and counts persist: the assign kernel maintains them incrementally
      if (A.bstate[2 * b] == 0 && !any_empty) A.bstate[2 * b + 1] = 1;
      A.bstate[2 * b] = 0;
    }
    __syncthreads();
  }
}

// Final assignment + probe list for the rows of tiled buckets, first pass: one THREAD per row,
// float32 scores against all (<= 32) lists of the bucket at once.  A block covers 256 consecutive
// rows; for every tiled bucket it touches it stages the transposed training copy of the centroids
// ([column][32 lists], zero padded) in shared memory, and each thread walks its sparse row once,
// reading the 32 scores' operands as eight float4.  Thread t reads them in the swizzled order
// r ^ (t & 7) (see lds128), so the eight threads of a quarter warp always hit eight different bank
// groups whatever their columns are; the accumulators stay in that order (static register
// indices) and are mapped back only when the best lists are named.
// The float32 order is final only where it is provably the float64 order: every gap between
// consecutive scores of the P + 1 best exceeds `gap` * sum|row values| (above twice the float32
// rounding bound of the sequential fmaf inner product with a unit-norm centroid).  All other rows, and the rows of buckets with more than 32 lists or
// more than 8 probes, are appended to `slow_rows` for the float64 kernel below.
constexpr int kAssignRows = 256;

__global__ void __launch_bounds__(kAssignRows, 3)
ivf_assign_tiled_f32_kernel(TiledArgs A, const int32_t* __restrict__ nprobe, int32_t max_nprobe, float gap,
                            int32_t* __restrict__ list_id, int32_t* __restrict__ probes,
                            int32_t* __restrict__ slow_rows, int32_t* __restrict__ slow_count) {
  extern __shared__ __align__(128) float4 cs4[];  // [d][8] float4 = [column][32 lists]
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int d = static_cast<int>(A.low_dim);
  const uint32_t cs_addr = static_cast<uint32_t>(__cvta_generic_to_shared(cs4));  // 128-byte aligned
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kAssignRows;
  const int64_t row1 = min(A.n, row0 + kAssignRows);
  if (row0 >= row1) return;
  int64_t b = find_segment(A.bucket_ptr, A.n_buckets, row0);
  // b < n_buckets: with a caller-side bound of the bucket count that turned out too small the bucket list does not
  // end at row n (the caller redoes the batch; this kernel must only stay inside the list)
  for (int64_t seg = row0; seg < row1 && b < A.n_buckets; ++b) {
    const int64_t seg_end = min(row1, A.bucket_ptr[b + 1]);
    const int64_t i = seg + tid;
    const bool active = i < seg_end;
    seg = seg_end;
    if (A.q.bclass[b] != kClsTiled) continue;
    const int32_t L = A.nlist[b];
    const int32_t P = min(min(nprobe[b], L), max_nprobe);
    bool slow = active;
    if (L <= 32 && P <= 8) {
      __syncthreads();  // the previous bucket's centroids are no longer read
      const float* ctb = A.ct + A.centroid_ptr[b] * d;
      if (L == 32) {
        const float4* src = reinterpret_cast<const float4*>(ctb);  // centroid_ptr * d * 4 bytes: 16-byte aligned when d % 4 == 0
        if ((reinterpret_cast<uintptr_t>(ctb) & 15) == 0) {
          for (int e = tid; e < d * 8; e += kAssignRows) cs4[e] = __ldg(src + e);
        } else {
          float* cs = reinterpret_cast<float*>(cs4);
          for (int e = tid; e < d * 32; e += kAssignRows) cs[e] = __ldg(ctb + e);
        }
      } else {
        float* cs = reinterpret_cast<float*>(cs4);
        for (int e = tid; e < d * 32; e += kAssignRows) {
          const int k = e >> 5, c = e & 31;
          cs[e] = c < L ? __ldg(ctb + static_cast<int64_t>(k) * L + c) : 0.f;
        }
      }
      __syncthreads();
      if (active) {
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        const int rot = tid & 7;
        float sabs = 0.f;  // sum |value|: scales the rounding bound (centroid entries are <= 1 in magnitude)
        const int m = min(static_cast<int>(A.ell_nnz[i]), A.W);
        const uint4* ip = reinterpret_cast<const uint4*>(A.ell_idx + i * A.W);   // W % 8 == 0: 16-byte aligned rows
        const float4* vp = reinterpret_cast<const float4*>(A.ell_val + i * A.W);
        for (int j0 = 0; j0 < m; j0 += 8) {
          const uint4 iw = __ldg(ip + (j0 >> 3));
          const float4 va = __ldg(vp + (j0 >> 2)), vb = __ldg(vp + (j0 >> 2) + 1);
          const uint32_t ks[8] = {iw.x & 0xffffu, iw.x >> 16, iw.y & 0xffffu, iw.y >> 16,
                                  iw.z & 0xffffu, iw.z >> 16, iw.w & 0xffffu, iw.w >> 16};
          const float vs[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            // slots past the row's population are zero padded (column 0, value 0): they add nothing
            const uint32_t rowa = cs_addr + ks[t] * 128u + static_cast<uint32_t>(rot) * 16u;
            const float v = vs[t];
            sabs += fabsf(v);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float4 c = lds128(rowa ^ (r * 16u));  // chunk r ^ rot
              acc[r][0] = fmaf(v, c.x, acc[r][0]);
              acc[r][1] = fmaf(v, c.y, acc[r][1]);
              acc[r][2] = fmaf(v, c.z, acc[r][2]);
              acc[r][3] = fmaf(v, c.w, acc[r][3]);
            }
          }
        }
        // register slot r*4+e holds list 4*(r ^ rot) + e; lists past L do not exist
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (4 * (r ^ rot) + e >= L) acc[r][e] = -INFINITY;
        const int T = min(P + 1, L);  // the runner-up after the last probe decides whether the cut is safe
        const float need = gap * sabs;
        float prev = 0.f;
        slow = false;
        for (int t = 0; t < T; ++t) {
          float bv = -INFINITY;
          int bs = 0;
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (acc[r][e] > bv) { bv = acc[r][e]; bs = r * 4 + e; }
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (bs == r * 4 + e) acc[r][e] = -INFINITY;
          if (t > 0 && !(prev - bv > need)) slow = true;
          prev = bv;
          if (t < P) {
            const int32_t id = 4 * ((bs >> 2) ^ rot) + (bs & 3);
            if (t < max_nprobe) probes[i * max_nprobe + t] = id;
            if (t == 0) list_id[i] = id;
          }
        }
        for (int t = P; t < max_nprobe; ++t) probes[i * max_nprobe + t] = -1;
      }
    }
    const uint32_t vote = __ballot_sync(0xffffffffu, slow);
    if (vote != 0u) {
      int32_t base = 0;
      if (lane == 0) base = atomicAdd(slow_count, __popc(vote));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (slow) slow_rows[base + __popc(vote & ((1u << lane) - 1u))] = static_cast<int32_t>(i);
    }
  }
}

// Second pass, for the rows the first pass could not decide: one warp per row,
// one LANE per list (strips of 32 lists): the row's entries are broadcast with
// shuffles, the transposed training copy of the centroids ([column][list]) makes
// the gathers coalesced, and there is no reduction across lanes.  float64 inner
// products, best-first insertion in list order (ties to the lower id) -- the
// same results as ivf_assign_kernel.
__global__ void __launch_bounds__(256)
ivf_assign_tiled_kernel(TiledArgs A, const int32_t* __restrict__ nprobe, int32_t max_nprobe,
                        const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows,
                        int32_t* __restrict__ list_id, int32_t* __restrict__ probes) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  const int lane = threadIdx.x & 31;
  unsigned char* pl_mem = pl_smem + (threadIdx.x >> 5) * probe_list_bytes(max_nprobe);
  const int32_t total_rows = *n_rows;
  for (int32_t w = blockIdx.x * 8 + (threadIdx.x >> 5); w < total_rows; w += gridDim.x * 8) {
  const int64_t i = rows[w];
  const int64_t b = find_segment(A.bucket_ptr, A.n_buckets, i);
  const int32_t L = A.nlist[b];
  const int32_t P = min(min(nprobe[b], L), max_nprobe);
  const int d = static_cast<int>(A.low_dim);
  const float* ctb = A.ct + A.centroid_ptr[b] * d;
  const int m = min(static_cast<int>(A.ell_nnz[i]), A.W);
  ProbeList top;
  top.init(P, reinterpret_cast<double*>(pl_mem), reinterpret_cast<int32_t*>(pl_mem + static_cast<size_t>(max_nprobe) * 8));
  for (int32_t s0 = 0; s0 < L; s0 += 32) {
    const int32_t c = s0 + lane;
    double acc = 0.0;
    for (int j0 = 0; j0 < m; j0 += 32) {
      const int j = j0 + lane;
      const uint32_t kj = j < m ? static_cast<uint32_t>(__ldg(A.ell_idx + i * A.W + j)) : 0u;
      const float vj = j < m ? __ldg(A.ell_val + i * A.W + j) : 0.f;
      const int cnt = min(32, m - j0);
      for (int t = 0; t < cnt; ++t) {
        const uint32_t k = __shfl_sync(0xffffffffu, kj, t);
        const float v = __shfl_sync(0xffffffffu, vj, t);
        if (c < L) acc = fma(static_cast<double>(v), static_cast<double>(__ldg(ctb + static_cast<int64_t>(k) * L + c)), acc);
      }
    }
    const int n_here = min(32, L - s0);
    if (L <= 32 && P <= 8) {
      // one strip, few probes: P rounds of warp arg-max (ties to the lower id) beat 32 insertions
      double mine = c < L ? acc : -INFINITY;
      bool taken = c >= L;
      for (int t = 0; t < P; ++t) {
        double bv = taken ? -INFINITY : mine;
        int bc = taken ? 0x7fffffff : static_cast<int>(c);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
          if (oc != 0x7fffffff && (bc == 0x7fffffff || ov > bv || (ov == bv && oc < bc))) { bv = ov; bc = oc; }
        }
        if (lane == t) { top.my_score = bv; top.my_id = bc; }
        if (static_cast<int>(c) == bc) taken = true;
      }
      break;
    }
    if (top.big && top.cnt == top.P) {
      // only the scores that beat the current P-th best need the list
      for (uint32_t rest = __ballot_sync(0xffffffffu, c < L && acc > top.worst); rest != 0u; rest &= rest - 1u) {
        const int t = __ffs(rest) - 1;
        top.push(__shfl_sync(0xffffffffu, acc, t), s0 + t, lane);
      }
    } else {
      for (int t = 0; t < n_here; ++t) top.push(__shfl_sync(0xffffffffu, acc, t), s0 + t, lane);
    }
  }
  top.store(probes + i * max_nprobe, max_nprobe, list_id + i, lane);
  }
}

struct KmeansLayout {
  int32_t* qctr;  // [8]: pull counters, queue lengths
  int32_t* queue;
  uint8_t* bclass;
  // tiled only
  float* ct;
  long long* gsum;
  int32_t* gcnt;
  double* gcntd;
  int32_t* gassign;
  int32_t* bstate;
  // tensor-core assignment
  uint16_t* cb;
  int32_t* tc_best;
  int32_t* tc_unsure;
  int4* units;
  int32_t* unit_bucket;
  int4* units_sp;
  int4* units_dn;
};

static void kmeans_layout(Workspace& ws, int64_t n, int64_t n_buckets, int64_t total, uint32_t low_dim,
                          bool tiled, KmeansLayout& L) {
  const size_t nbk = static_cast<size_t>(n_buckets > 0 ? n_buckets : 1);
  L.qctr = ws.take<int32_t>(16);  // [0,4) pull counters, [4,8) queue lengths, [8,12) tensor-core counts, [12] negative value seen
  L.queue = ws.take<int32_t>(3 * nbk);
  L.bclass = ws.take<uint8_t>(nbk);
  L.ct = nullptr; L.gsum = nullptr; L.gcnt = nullptr; L.gcntd = nullptr; L.gassign = nullptr; L.bstate = nullptr;
  L.cb = nullptr; L.tc_best = nullptr; L.tc_unsure = nullptr; L.units = nullptr; L.unit_bucket = nullptr;
  L.units_sp = nullptr; L.units_dn = nullptr;
  if (tiled) {
    const size_t t = static_cast<size_t>(total > 0 ? total : 1);
    L.ct = ws.take<float>(t * low_dim);
    L.gsum = ws.take<long long>(t * low_dim);
    L.gcnt = ws.take<int32_t>(t);
    L.gcntd = ws.take<double>(t);
    L.gassign = ws.take<int32_t>(static_cast<size_t>(n > 0 ? n : 1));
    L.bstate = ws.take<int32_t>(2 * nbk);
    const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
    L.cb = ws.take<uint16_t>(t * ((low_dim + 7u) & ~7u));
    L.tc_best = ws.take<int32_t>(nn);
    L.tc_unsure = ws.take<int32_t>(nn);
    L.units = ws.take<int4>(nn / 128 + nbk + 1);
    L.unit_bucket = ws.take<int32_t>(nn / 128 + nbk + 1);
    L.units_sp = ws.take<int4>(nn / 128 + nbk + 1);
    L.units_dn = ws.take<int4>(nn / 128 + nbk + 1);
  }
}

// Can any bucket be too large for the fused trainer?  Upper bound: every ELL slot
// of the largest IVF bucket populated.
static bool kmeans_needs_tiled(int64_t n, int64_t max_ivf_bucket, int32_t W, uint32_t low_dim) {
  const int64_t nb = max_ivf_bucket > 0 ? max_ivf_bucket : n;
  if (nb > 32767) return true;
  const int32_t L = nlist_rule(nb);
  if (L > kFusedMaxL) return true;
  return fused_plan(static_cast<int>(nb), W | 1, L, static_cast<int>(low_dim), fused_threads(1) / 32).total > fused_smem_limit(1);
}

}  // namespace flc

extern "C" {

int flc_ivf_plan(const int64_t* bucket_ptr, int64_t n_buckets, int32_t n_probe, int exhaustive,
                 int32_t* nlist, int32_t* nprobe, int64_t* centroid_ptr, int64_t* total_centroids,
                 int32_t* max_nprobe, int64_t* max_ivf_bucket, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n_buckets >= 0, "n_buckets must be non-negative");
  FLC_REQUIRE(n_probe >= 1, "n_probe must be >= 1");
  FLC_REQUIRE((total_centroids == nullptr) == (max_nprobe == nullptr), "total_centroids and max_nprobe go together");
  cudaStream_t stream = as_stream(stream_);
  timed("ivf_plan", stream, [&] { ivf_plan_kernel<<<1, 1024, 0, stream>>>(
      bucket_ptr, n_buckets, n_probe, exhaustive, nlist, nprobe, centroid_ptr); });
  FLC_LAUNCH_CHECK();
  if (total_centroids == nullptr) return FLC_OK;  // no synchronisation: the totals stay in centroid_ptr[n_buckets ..]
  int64_t tail[3] = {0, 0, 0};
  FLC_CUDA(cudaMemcpyAsync(tail, centroid_ptr + n_buckets, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  FLC_CUDA(cudaStreamSynchronize(stream));
  *total_centroids = tail[0];
  *max_nprobe = static_cast<int32_t>(tail[1] > 0 ? tail[1] : 1);
  if (max_ivf_bucket) *max_ivf_bucket = tail[2];
  return FLC_OK;
}

size_t flc_kmeans_workspace_bytes(int64_t n, int64_t n_buckets, int64_t total_centroids,
                                  int64_t max_ivf_bucket, int32_t ell_width, uint32_t low_dim) {
  flc::Workspace ws(nullptr, 0);
  flc::KmeansLayout L;
  const char* force_env = getenv("FLC_KMEANS_FORCE_TILED");
  const bool force_tiled = force_env != nullptr && force_env[0] == '1';
  flc::kmeans_layout(ws, n, n_buckets, total_centroids, low_dim,
                     force_tiled || flc::kmeans_needs_tiled(n, max_ivf_bucket, ell_width, low_dim), L);
  return ws.used + 256;
}

int flc_kmeans_needs_tiled(int64_t n, int64_t max_ivf_bucket, int32_t ell_width, uint32_t low_dim) {
  const char* force_env = getenv("FLC_KMEANS_FORCE_TILED");
  if (force_env != nullptr && force_env[0] == '1') return 1;
  return flc::kmeans_needs_tiled(n, max_ivf_bucket, ell_width, low_dim) ? 1 : 0;
}

int flc_kmeans_train(const uint16_t* ell_idx, const float* ell_val, const uint16_t* ell_nnz, int32_t ell_width,
                     const uint16_t* x_bf16, int64_t ld_bf16,
                     int64_t n, uint32_t low_dim, const int64_t* bucket_ptr, int64_t n_buckets,
                     const int32_t* nlist, const int64_t* centroid_ptr, int64_t total_centroids,
                     int64_t max_ivf_bucket, int niter, float* centroids, const int32_t* nprobe,
                     int32_t max_nprobe, int32_t* list_id, int32_t* probes, void* workspace,
                     size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && niter >= 0 && n_buckets >= 0, "bad sizes");
  FLC_REQUIRE(low_dim > 0 && low_dim <= 8192, "low_dim must be in [1, 8192]");
  FLC_REQUIRE(ell_idx && ell_val && ell_nnz, "k-means trains on the sparse rows of flc_vectorize");
  FLC_REQUIRE(ell_width > 0 && (ell_width % 8) == 0, "ell_width must be a positive multiple of 8");
  FLC_REQUIRE((reinterpret_cast<uintptr_t>(ell_idx) % 16) == 0 && (reinterpret_cast<uintptr_t>(ell_val) % 16) == 0,
              "ELL arrays must be 16-byte aligned");
  FLC_REQUIRE((list_id == nullptr) == (probes == nullptr), "list_id and probes go together");
  FLC_REQUIRE(list_id == nullptr || (nprobe != nullptr && max_nprobe >= 1), "probe outputs need nprobe / max_nprobe");
  if (n == 0 || n_buckets == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
  const int32_t W = ell_width;
  // FLC_KMEANS_FORCE_TILED=1 sends every bucket through the tiled trainer (tests
  // use it to check that the two schedules agree bit for bit).
  const char* force_env = getenv("FLC_KMEANS_FORCE_TILED");
  const int force_tiled = (force_env != nullptr && force_env[0] == '1') ? 1 : 0;
  const bool tiled = force_tiled || kmeans_needs_tiled(n, max_ivf_bucket, W, low_dim);
  FLC_REQUIRE(max_nprobe <= 1024, "n_probe > 1024 not supported");
  Workspace ws(workspace, workspace_bytes);
  KmeansLayout K;
  kmeans_layout(ws, n, n_buckets, total_centroids, low_dim, tiled, K);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "kmeans workspace too small: need %zu", ws.used);
  TrainQueues q{K.qctr, K.qctr + 4, K.queue, K.bclass};
  FLC_CUDA(cudaMemsetAsync(K.qctr, 0, 16 * sizeof(int32_t), stream));
  timed("kmeans_classify", stream, [&] {
    kmeans_classify_kernel<<<static_cast<unsigned>((n_buckets + 7) / 8), 256, 0, stream>>>(
        ell_nnz, low_dim, bucket_ptr, n_buckets, nlist, q, force_tiled, max_nprobe, list_id, probes, centroid_ptr,
        total_centroids); });
  FLC_LAUNCH_CHECK();
  if (total_centroids == 0) return FLC_OK;
  // ---- fused classes
  unsigned long long* timing = nullptr;
  const char* timing_env = getenv("FLC_KMEANS_TIMING");
  if (timing_env != nullptr && timing_env[0] == '1')
    FLC_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&timing), g_kmeans_timing));
  FusedArgs A{ell_idx, ell_val, ell_nnz, W, low_dim, bucket_ptr, n_buckets, nlist, centroid_ptr, niter, q,
              centroids, nprobe, max_nprobe, list_id, probes, timing};
  {
    FLC_CUDA(cudaFuncSetAttribute(kmeans_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(fused_smem_limit(0))));
    const int64_t grid0 = std::min<int64_t>(n_buckets, 2 * kNumSMs);
    timed("kmeans_fused", stream, [&] {
      kmeans_fused_kernel<256><<<static_cast<unsigned>(grid0), 256, fused_smem_limit(0), stream>>>(A, 0); });
    FLC_LAUNCH_CHECK();
    FLC_CUDA(cudaFuncSetAttribute(kmeans_fused_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(fused_smem_limit(1))));
    const int64_t grid1 = std::min<int64_t>(n_buckets, kNumSMs);
    timed("kmeans_fused_large", stream, [&] {
      kmeans_fused_kernel<512><<<static_cast<unsigned>(grid1), 512, fused_smem_limit(1), stream>>>(A, 1); });
    FLC_LAUNCH_CHECK();
  }
  if (!tiled) return FLC_OK;
  // ---- tiled trainer for what is left
  // Tensor-core assignment needs the bf16 rows of unit-norm vectors (what flc_vectorize emits with norm = 1);
  // FLC_KMEANS_NO_TC=1 keeps the SIMT assignment (tests compare the two).
  const char* no_tc_env = getenv("FLC_KMEANS_NO_TC");
  const bool use_tc = x_bf16 != nullptr && !(no_tc_env != nullptr && no_tc_env[0] == '1') &&
                      (ld_bf16 % 8) == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0;
  const int64_t ld_c = (static_cast<int64_t>(low_dim) + 7) & ~int64_t(7);
  TiledArgs T{ell_idx, ell_val, ell_nnz, W, low_dim, n, bucket_ptr, n_buckets, nlist, centroid_ptr, q,
              centroids, K.ct, K.gsum, K.gcnt, K.gcntd, K.gassign, K.bstate,
              use_tc ? K.cb : nullptr, ld_c, K.tc_best, K.tc_unsure, K.units, K.unit_bucket, K.qctr + 8,
              K.units_sp, K.units_dn};
  FLC_CUDA(cudaMemsetAsync(K.gsum, 0, static_cast<size_t>(total_centroids) * low_dim * sizeof(long long), stream));
  FLC_CUDA(cudaMemsetAsync(K.gcnt, 0, static_cast<size_t>(total_centroids) * sizeof(int32_t), stream));
  FLC_CUDA(cudaMemsetAsync(K.gassign, 0xff, static_cast<size_t>(n) * sizeof(int32_t), stream));
  FLC_CUDA(cudaMemsetAsync(K.bstate, 0, static_cast<size_t>(n_buckets) * 2 * sizeof(int32_t), stream));
  timed("kmeans_tiled_init", stream, [&] {
    kmeans_tiled_init_kernel<<<static_cast<unsigned>(total_centroids), 128, 0, stream>>>(T, total_centroids); });
  FLC_LAUNCH_CHECK();
  const size_t smem = static_cast<size_t>(low_dim) * kTiledG * sizeof(float);
  FLC_REQUIRE(smem <= 200 * 1024, "low_dim too large for the tiled trainer");
  FLC_CUDA(cudaFuncSetAttribute(kmeans_tiled_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  const unsigned row_blocks = static_cast<unsigned>((n + kTiledRows - 1) / kTiledRows);
  static_assert(kTiledThreads == 256, "launch configuration below assumes 256 threads");
  const unsigned upd_blocks = static_cast<unsigned>(std::min<int64_t>(n_buckets, 1 << 20));  // latency-bound: one CTA per bucket
  const unsigned unit_blocks = static_cast<unsigned>(std::min<int64_t>(n / 128 + n_buckets + 1, 1 << 20));
  const size_t upd_smem = 32 * (static_cast<size_t>(low_dim) + 1) * sizeof(float);
  FLC_REQUIRE(upd_smem <= 200 * 1024, "low_dim too large for the tiled trainer");
  FLC_CUDA(cudaFuncSetAttribute(kmeans_tiled_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(upd_smem)));
  // Who assigns: with the bf16 rows, every tiled bucket goes through the tensor cores (832 bytes a row, HBM
  // bound: ~0.17 ms per million rows and iteration on B200, plus the exact re-score of the close calls).  The
  // thread-per-row kernel reads the sparse rows (~200 bytes) but is bound by the shared-memory gathers at
  // ~0.30 ms, so it is the fallback when there are no bf16 rows.  FLC_KMEANS_SIMT_SMALL=1 gives it the buckets
  // of up to 32 lists (tests compare the schedules).
  const char* simt_env = getenv("FLC_KMEANS_SIMT_SMALL");
  const bool tc_all = use_tc && !(simt_env != nullptr && simt_env[0] == '1');
  const bool tc_some = use_tc;
  // FLC_KMEANS_TC_DENSE=1: every tensor-core tile reads the dense bf16 rows (tests compare the two kernels)
  const char* dense_env = getenv("FLC_KMEANS_TC_DENSE");
  const bool sparse_tc = use_tc && kmeans_tc_sparse_ok(low_dim, W) && !(dense_env != nullptr && dense_env[0] == '1');
  const int32_t simt_max_lists = tc_all ? 0 : (tc_some ? kTiledG : 0x7fffffff);
  for (int it = 0; it < niter; ++it) {
    if (simt_max_lists > 0) {
      timed("kmeans_tiled_assign", stream, [&] {
        kmeans_tiled_assign_kernel<<<row_blocks, kTiledThreads, smem, stream>>>(T, simt_max_lists); });
      FLC_LAUNCH_CHECK();
    }
    if (tc_some) {
      // query tiles of the buckets that have not converged yet
      FLC_CUDA(cudaMemsetAsync(T.tc_counts, 0, 4 * sizeof(int32_t), stream));
      timed("kmeans_tc_units", stream, [&] {
        kmeans_tc_units_kernel<<<static_cast<unsigned>((n_buckets + 7) / 8), 256, 0, stream>>>(
            T, tc_all ? 0 : kTiledG + 1, sparse_tc ? kSparseMaxLists : 0); });
      FLC_LAUNCH_CHECK();
      // bf16 scores on the tensor cores decide every row whose two best lists are further apart than
      // twice the rounding error (2^-7 for unit vectors, plus slack); the rest is flagged and re-scored exactly
      // by the apply kernel.
      // Buckets of up to 64 lists are scored from the sparse rows (expanded in shared memory), the others
      // from the dense bf16 rows.
      if (sparse_tc)
        FLC_TRY(launch_kmeans_tc_sparse(ell_idx, ell_val, ell_nnz, W, K.cb, ld_c, total_centroids, low_dim, K.units_sp,
                                        T.tc_counts + 2, 0.008f, K.qctr + 12, it > 0 ? 1 : 0, K.tc_best, stream));
      FLC_TRY(launch_kmeans_tc(x_bf16, ld_bf16, n, K.cb, ld_c, total_centroids, low_dim, K.units_dn, T.tc_counts + 3,
                               0.008f, K.tc_best, stream));
      timed("kmeans_tiled_apply", stream, [&] { kmeans_tiled_apply_kernel<<<unit_blocks, 128, 0, stream>>>(T); });
      FLC_LAUNCH_CHECK();
    }
    timed("kmeans_tiled_update", stream, [&] { kmeans_tiled_update_kernel<<<upd_blocks, 256, upd_smem, stream>>>(T); });
    FLC_LAUNCH_CHECK();
  }
  if (list_id != nullptr) {
    // float32 pass for every row; what it cannot decide goes through the float64 kernel
    const size_t asmem = static_cast<size_t>(low_dim) * 32 * sizeof(float);
    FLC_REQUIRE(asmem <= 200 * 1024, "low_dim too large for the tiled assignment");
    FLC_CUDA(cudaFuncSetAttribute(ivf_assign_tiled_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(asmem)));
    FLC_CUDA(cudaMemsetAsync(T.tc_counts + 1, 0, sizeof(int32_t), stream));
    // |float32 - float64| of W sequential fmaf <= W * 2^-24 * sum|v_j c_j| <= W * 2^-24 * sum|v_j|; the order of
    // two scores is safe when they differ by more than twice that.  2x slack on top: (W + 8) * 2^-22 * sum|v_j|
    const float gap = static_cast<float>(W + 8) * 2.384185791015625e-7f;
    timed("ivf_assign_tiled_f32", stream, [&] {
      ivf_assign_tiled_f32_kernel<<<static_cast<unsigned>((n + kAssignRows - 1) / kAssignRows), kAssignRows, asmem, stream>>>(
          T, nprobe, max_nprobe, gap, list_id, probes, K.tc_unsure, T.tc_counts + 1); });
    FLC_LAUNCH_CHECK();
    const size_t plsmem = 8 * probe_list_bytes(max_nprobe);
    if (plsmem > 48 * 1024)
      FLC_CUDA(cudaFuncSetAttribute(ivf_assign_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(plsmem)));
    timed("ivf_assign_tiled", stream, [&] { ivf_assign_tiled_kernel<<<kNumSMs * 8, 256, plsmem, stream>>>(
        T, nprobe, max_nprobe, K.tc_unsure, T.tc_counts + 1, list_id, probes); });
    FLC_LAUNCH_CHECK();
  }
  return FLC_OK;
}

/* Debug aid: copies out and clears the 16 per-phase cycle counters of the fused trainer
 * (filled when FLC_KMEANS_TIMING=1; see g_kmeans_timing). */
int flc_debug_kmeans_timing(unsigned long long* out16) {
  using namespace flc;
  FLC_REQUIRE(out16 != nullptr, "null output");
  FLC_CUDA(cudaDeviceSynchronize());
  FLC_CUDA(cudaMemcpyFromSymbol(out16, g_kmeans_timing, 16 * sizeof(unsigned long long)));
  const unsigned long long zero[16] = {0};
  FLC_CUDA(cudaMemcpyToSymbol(g_kmeans_timing, zero, sizeof(zero)));
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
  FLC_REQUIRE(max_nprobe <= 1024, "n_probe > 1024 not supported");
  if (n == 0) return FLC_OK;
  cudaStream_t stream = as_stream(stream_);
  const size_t smem = static_cast<size_t>(8) * low_dim * sizeof(float) + 8 * probe_list_bytes(max_nprobe);
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(ivf_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  timed("ivf_assign", stream, [&] { ivf_assign_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, smem, stream>>>(
      x, ld, n, low_dim, bucket_ptr, n_buckets, nlist, nprobe, centroid_ptr, centroids, max_nprobe, ell_idx,
      ell_val, ell_width, nullptr, list_id, probes); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
