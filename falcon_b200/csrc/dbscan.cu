// Stage 4: DBSCAN on the sparse k-NN matrix + precursor-tolerance split.
//
// DBSCAN (A.4): neighbourhood(i) = CSR row entries with dist <= eps, core(i) =
// |neighbourhood| >= min_samples.  sklearn's dbscan_inner
// (sklearn/cluster/_dbscan_inner.pyx:11-41) visits points in index order and
// grows clusters along edges that LEAVE core points, so on this asymmetric
// k-NN graph a point's cluster is the rank of the minimum-index core point it
// is reachable from -- not an undirected component.  The
// kernels compute that fix-point: m[core u] = u, every core u pushes
// atomicMin(m[w], m[u]) along its out-edges, chains are shortcut by pointer
// jumping (m[u] is itself a core point that reaches u, so m[m[u]] reaches u),
// repeat until nothing changes; labels = rank of m[v] among the seeds
// (m[s] == s) by prefix sum.
//
// Precursor split (falcon/cluster/cluster.py:334-509): inside each DBSCAN
// cluster the members, ascending in precursor m/z, are agglomerated with 1-D
// complete linkage and the dendrogram is cut at the tolerance.  The reference
// merges the globally closest adjacent pair first; because complete linkage is
// reducible, merging every adjacent pair that is a local minimum of the
// merge distance (strict on the left, non-strict on the right = "first minimum
// wins") in parallel rounds yields the same flat clusters.  One warp per
// cluster keeps the list of run starts in global scratch and compacts it each
// round with ballots.
//
// HBM/latency-bound: nnz * 8 bytes for the eps mask pass, then
// sweeps * (nnz_eps * 4 + N * 8), one 32-bit key sort of N rows.
#include <cub/cub.cuh>

#include "common.cuh"

namespace flc {

constexpr int32_t kInf = 0x7fffffff;
constexpr uint32_t kNoiseKey = 0xffffffffu;

// ---------------------------------------------------------------- DBSCAN
// One thread per row: rows hold at most n_neighbors (64) entries, a handful on average.
__global__ void dbscan_core_kernel(const float* __restrict__ dist, const int64_t* __restrict__ indptr,
                                   int64_t n, float eps, int32_t min_samples, int32_t* __restrict__ m,
                                   uint8_t* __restrict__ core) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t a = indptr[i], b = indptr[i + 1];
  int cnt = 0;
  for (int64_t p = a; p < b; ++p) cnt += (__ldg(dist + p) <= eps) ? 1 : 0;
  const bool c = cnt >= min_samples;
  core[i] = c ? 1 : 0;
  m[i] = c ? static_cast<int32_t>(i) : kInf;
}

__global__ void dbscan_propagate_kernel(const float* __restrict__ dist, const int32_t* __restrict__ indices,
                                        const int64_t* __restrict__ indptr, int64_t n, float eps,
                                        const uint8_t* __restrict__ core, int32_t* m,
                                        int32_t* __restrict__ changed) {
  const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (u >= n || !core[u]) return;
  const int64_t a = indptr[u], b = indptr[u + 1];
  bool any = false;
  // Push until this row's own value stops improving under us: neighbours run the
  // same loop concurrently, so small components settle within a single sweep.
  for (int round = 0; round < 64; ++round) {
    // Shortcut: follow the chain of minimum ancestors to its current end.
    int32_t mu = *reinterpret_cast<volatile int32_t*>(m + u);
    int32_t r = mu;
    while (true) {
      const int32_t next = *reinterpret_cast<volatile int32_t*>(m + r);
      if (next >= r) break;
      r = next;
    }
    if (r < mu) {
      atomicMin(m + u, r);
      any = true;
      mu = r;
    }
    for (int64_t p = a; p < b; ++p) {
      if (__ldg(dist + p) <= eps) {
        const int32_t w = __ldg(indices + p);
        if (*reinterpret_cast<volatile int32_t*>(m + w) > mu) {
          const int32_t old = atomicMin(m + w, mu);
          any |= old > mu;
        }
      }
    }
    if (*reinterpret_cast<volatile int32_t*>(m + u) >= mu) break;
  }
  if (any) *changed = 1;
}

__global__ void dbscan_seed_kernel(const int32_t* __restrict__ m, const uint8_t* __restrict__ core, int64_t n,
                                   int32_t* __restrict__ seed) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i <= n) seed[i] = (i < n && core[i] && m[i] == static_cast<int32_t>(i)) ? 1 : 0;
}

__global__ void dbscan_label_kernel(const int32_t* __restrict__ m, const int32_t* __restrict__ rank, int64_t n,
                                    int32_t* __restrict__ labels, int64_t* __restrict__ n_clusters_dev) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == n && n_clusters_dev != nullptr) *n_clusters_dev = rank[n];  // the grid covers n + 1 threads
  if (i < n) {
    const int32_t r = m[i];
    labels[i] = r == kInf ? -1 : rank[r];
  }
}

struct DbscanLayout {
  int32_t* m;
  uint8_t* core;
  int32_t* seed;
  int32_t* rank;
  int32_t* changed;
  void* cub_tmp;
  size_t cub_bytes;
};

static void dbscan_layout(Workspace& ws, int64_t n, DbscanLayout& L) {
  L.m = ws.take<int32_t>(n + 1);
  L.core = ws.take<uint8_t>(n + 1);
  L.seed = ws.take<int32_t>(n + 1);
  L.rank = ws.take<int32_t>(n + 1);
  L.changed = ws.take<int32_t>(4);
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, static_cast<int>(n + 1));
  L.cub_bytes = b;
  L.cub_tmp = ws.take<char>(b);
}

// ---------------------------------------------------------------- precursor split
__global__ void split_key_kernel(const int32_t* __restrict__ labels, int64_t n, uint32_t* __restrict__ key,
                                 int32_t* __restrict__ idx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const int32_t l = labels[i];
    key[i] = l < 0 ? kNoiseKey : static_cast<uint32_t>(l);
    idx[i] = static_cast<int32_t>(i);
  }
}

__global__ void split_mzkey_kernel(const double* __restrict__ mz, int64_t n, uint64_t* __restrict__ key,
                                   int32_t* __restrict__ idx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t b = static_cast<uint64_t>(__double_as_longlong(mz[i]));
    key[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    idx[i] = static_cast<int32_t>(i);
  }
}

// After sorting by label: gather the values, flag group heads (label change).
__global__ void split_prepare_kernel(const uint32_t* __restrict__ key_sorted, const int32_t* __restrict__ perm,
                                     const double* __restrict__ mz, int64_t n, double* __restrict__ vs,
                                     uint8_t* __restrict__ ghead, uint8_t* __restrict__ runhead) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  vs[i] = mz[perm[i]];
  const uint32_t k = key_sorted[i];
  const bool head = (i == 0) || (k != key_sorted[i - 1]);
  ghead[i] = head ? 1 : 0;
  // Noise rows are runs of their own; cluster members start out as heads only at
  // the group start, the split kernel adds the interior run starts.
  runhead[i] = (head || k == kNoiseKey) ? 1 : 0;
}

__device__ __forceinline__ double split_distance(double lo, double hi, int tol_mode) {
  const double d = hi - lo;
  return tol_mode == FLC_TOL_PPM ? d / lo * 1000000.0 : d;
}

// One warp per DBSCAN cluster.  list_a / list_b: per-element scratch holding
// the current run starts (offsets inside the group), ping-pong.
__device__ void split_group_one(int64_t g, int64_t n_groups, const int64_t* __restrict__ gstart,
                                const uint32_t* __restrict__ key_sorted, const double* __restrict__ vs,
                                int64_t n, double tol, int tol_mode, int32_t* list_a, int32_t* list_b,
                                uint8_t* __restrict__ runhead) {
  const int lane = threadIdx.x & 31;
  const int64_t s = gstart[g];
  const int64_t e = (g + 1 < n_groups) ? gstart[g + 1] : n;
  if (key_sorted[s] == kNoiseKey) return;
  const int32_t m = static_cast<int32_t>(e - s);
  if (m < 2) return;
  const double* v = vs + s;
  if (m <= 32) {
    // Whole group in one warp: lane j holds element j, the run starts are a bit mask.
    // Boundary j separates the run ending at j - 1 from the run starting at j; its
    // merge distance is the span of both runs.  d of the neighbouring boundaries is
    // what the lanes at the previous / next run start computed.
    const double val = lane < m ? v[lane] : 0.0;
    uint32_t heads = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
    while (true) {
      const bool is_b = lane >= 1 && lane < m && ((heads >> lane) & 1u);
      const uint32_t lower = heads & ((1u << lane) - 1u);
      const int ps = lower ? 31 - __clz(lower) : 0;                       // start of the run before the boundary
      const uint32_t above = lane < 31 ? (heads >> (lane + 1)) : 0u;
      const int ne = above ? min(lane + __ffs(above), m) : m;             // end of the run after the boundary
      const double d = split_distance(__shfl_sync(0xffffffffu, val, ps), __shfl_sync(0xffffffffu, val, ne - 1), tol_mode);
      const double dl = __shfl_sync(0xffffffffu, d, ps);
      const double dr = __shfl_sync(0xffffffffu, d, ne < m ? ne : 0);
      bool merge = is_b && d <= tol;
      if (merge && ps >= 1 && !(d < dl)) merge = false;
      if (merge && ne < m && !(d <= dr)) merge = false;
      const uint32_t mb = __ballot_sync(0xffffffffu, merge);
      if (mb == 0u) break;
      heads &= ~mb;
    }
    if (lane < m && ((heads >> lane) & 1u)) runhead[s + lane] = 1;
    return;
  }
  int32_t* cur = list_a + s;
  int32_t* nxt = list_b + s;
  for (int32_t j = lane; j < m; j += 32) cur[j] = j;
  __syncwarp();
  int32_t r = m;
  // boundary j (1 <= j < r) separates run j-1 = [cur[j-1], cur[j]) from run j.
  auto run_start = [&](int32_t j) -> int32_t { return __ldcg(cur + j); };
  auto run_end = [&](int32_t j) -> int32_t { return j + 1 < r ? __ldcg(cur + j + 1) : m; };
  while (r > 1) {
    int32_t r_new = 0;
    bool merged_any = false;
    for (int32_t base = 0; base < r; base += 32) {
      const int32_t j = base + lane;
      const bool valid = j < r;
      bool merge = false;
      if (valid && j >= 1) {
        const double d = split_distance(v[run_start(j - 1)], v[run_end(j) - 1], tol_mode);
        if (d <= tol) {
          merge = true;
          if (j >= 2) {
            const double dl = split_distance(v[run_start(j - 2)], v[run_end(j - 1) - 1], tol_mode);
            if (!(d < dl)) merge = false;
          }
          if (merge && j + 1 < r) {
            const double dr = split_distance(v[run_start(j)], v[run_end(j + 1) - 1], tol_mode);
            if (!(d <= dr)) merge = false;
          }
        }
      }
      const bool keep = valid && !merge;
      const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcg(nxt + r_new + __popc(ballot & ((1u << lane) - 1u)), run_start(j));
      r_new += __popc(ballot);
      merged_any |= __any_sync(0xffffffffu, merge);
    }
    __syncwarp();
    if (!merged_any) break;
    int32_t* t = cur; cur = nxt; nxt = t;
    r = r_new;
  }
  for (int32_t j = lane; j < r; j += 32) runhead[s + __ldcg(cur + j)] = 1;
}

// Four groups of at most eight members side by side in one warp (eight lanes each): the register
// algorithm of split_group_one at width 8.  Most DBSCAN clusters are this small.
__device__ __forceinline__ void split_groups_quad(int64_t s, int32_t m, bool active, const double* __restrict__ vs,
                                                  double tol, int tol_mode, uint8_t* __restrict__ runhead) {
  const int lane = threadIdx.x & 31;
  const int sl = lane & 7, base = lane & ~7;
  const double val = (active && sl < m) ? vs[s + sl] : 0.0;
  uint32_t heads = active ? ((1u << m) - 1u) : 0u;
  while (true) {
    const bool is_b = active && sl >= 1 && sl < m && ((heads >> sl) & 1u);
    const uint32_t lower = heads & ((1u << sl) - 1u);
    const int ps = lower ? 31 - __clz(lower) : 0;
    const uint32_t above = heads >> (sl + 1);
    const int ne = above ? min(sl + __ffs(above), m) : m;
    const double d = split_distance(__shfl_sync(0xffffffffu, val, base + ps),
                                    __shfl_sync(0xffffffffu, val, base + max(ne - 1, 0)), tol_mode);
    const double dl = __shfl_sync(0xffffffffu, d, base + ps);
    const double dr = __shfl_sync(0xffffffffu, d, base + (ne < m ? ne : 0));
    bool merge = is_b && d <= tol;
    if (merge && ps >= 1 && !(d < dl)) merge = false;
    if (merge && ne < m && !(d <= dr)) merge = false;
    const uint32_t mb = __ballot_sync(0xffffffffu, merge);
    if (mb == 0u) break;
    heads &= ~((mb >> base) & 0xffu);
  }
  if (active && sl < m && ((heads >> sl) & 1u)) runhead[s + sl] = 1;
}

// Warps stride over the groups, four at a time (their number is only known on the device).
__global__ void __launch_bounds__(256)
split_group_kernel(const int64_t* __restrict__ gstart, const int64_t* __restrict__ n_groups_ptr,
                   const uint32_t* __restrict__ key_sorted, const double* __restrict__ vs, int64_t n, double tol,
                   int tol_mode, int32_t* list_a, int32_t* list_b, uint8_t* __restrict__ runhead) {
  const int64_t n_groups = *n_groups_ptr;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t g0 = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * 4; g0 < n_groups;
       g0 += warps_total * 4) {
    const int64_t g = g0 + (lane >> 3);
    int64_t s = 0;
    int32_t m = 0;
    bool active = false;
    if (g < n_groups) {
      s = gstart[g];
      const int64_t e = (g + 1 < n_groups) ? gstart[g + 1] : n;
      const bool noise = key_sorted[s] == kNoiseKey;
      m = noise ? 0 : static_cast<int32_t>(min(e - s, static_cast<int64_t>(1 << 30)));
      active = m >= 2;
    }
    if (__all_sync(0xffffffffu, m <= 8)) {
      split_groups_quad(s, m, active, vs, tol, tol_mode, runhead);
    } else {
      for (int t = 0; t < 4; ++t) {
        if (g0 + t < n_groups) split_group_one(g0 + t, n_groups, gstart, key_sorted, vs, n, tol, tol_mode, list_a, list_b, runhead);
        __syncwarp();
      }
    }
    __syncwarp();
  }
}

// run id of every element = inclusive scan of runhead - 1; run start positions
// compacted into rstart; a run is kept when it has >= min_samples members and is
// not noise.
__global__ void split_keep_kernel(const int64_t* __restrict__ rstart, const int64_t* __restrict__ n_runs_ptr,
                                  const uint32_t* __restrict__ key_sorted, int64_t n, int32_t min_samples,
                                  int32_t* __restrict__ kept) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n_runs = *n_runs_ptr;
  if (r > n) return;
  if (r >= n_runs) {
    kept[r] = 0;
    return;
  }
  const int64_t a = rstart[r];
  const int64_t b = (r + 1 < n_runs) ? rstart[r + 1] : n;
  kept[r] = (key_sorted[a] != kNoiseKey && (b - a) >= min_samples) ? 1 : 0;
}

__global__ void split_label_kernel(const int32_t* __restrict__ run_id_incl, const int32_t* __restrict__ kept,
                                   const int32_t* __restrict__ new_id, const int32_t* __restrict__ perm,
                                   int64_t n, int32_t* __restrict__ labels_out, int64_t* __restrict__ n_clusters_dev) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0 && n_clusters_dev != nullptr) *n_clusters_dev = new_id[n];
  if (i >= n) return;
  const int32_t r = run_id_incl[i] - 1;
  labels_out[perm[i]] = kept[r] ? new_id[r] : -1;
}

struct U8ToI32 {
  __host__ __device__ __forceinline__ int32_t operator()(uint8_t v) const { return static_cast<int32_t>(v); }
};


// ---------------------------------------------------------------- RT cut (cluster.py:418-429)
// With rt_tol the reference cuts a second 1-D complete linkage (on retention time) and
// combines the two flat assignments as np.unique(a_mz * 2 + a_rt * 3).  That map is not
// injective, so the result depends on the ids scipy's fcluster hands out: a depth-first
// walk of the dendrogram from the root (scipy/cluster/_hierarchy.pyx cluster_monocrit) --
// left subtree, right subtree, then the children that are single observations, left before
// right; a subtree at or below the cut gets one id when it is entered.  So the whole
// dendrogram above the cut matters.  linkage_ids_one builds it with the same parallel
// local-minimum rounds as the split (first without crossing the tolerance = the flat
// clusters, then to the root) and turns the walk into range additions: merging nodes A
// (left) and B (right) shifts the ids of A's flat clusters by off(A) and B's by off(B),
//   A, B both single observations: 0, 1      A single only: |B|, 0
//   B single only: 0, |A|                    neither: 0, |A|
// (|X| = number of flat clusters in X); a flat cluster's id is the sum of the shifts on its
// path to the root = prefix sum of a difference array over the flat clusters, which are
// contiguous in sorted order.
__device__ void linkage_ids_one(int64_t g, int64_t n_groups, const int64_t* __restrict__ gstart,
                                const uint32_t* __restrict__ key_sorted, const double* __restrict__ vs,
                                int64_t n, double tol, int tol_mode, int32_t* list_a, int32_t* list_b,
                                int32_t* frank, int32_t* delta, int32_t* __restrict__ aid) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t s = gstart[g];
  const int64_t e = (g + 1 < n_groups) ? gstart[g + 1] : n;
  if (key_sorted[s] == kNoiseKey) return;
  const int32_t m = static_cast<int32_t>(e - s);
  if (m < 2) {
    if (lane == 0) aid[s] = 0;
    return;
  }
  const double* v = vs + s;
  int32_t* cur = list_a + s;
  int32_t* nxt = list_b + s;
  int32_t* fr = frank + s;
  int32_t* dl_ = delta + s + g;  // r0 + 1 <= m + 1 entries per group
  for (int32_t j = lane; j < m; j += 32) {
    __stcg(cur + j, j);
    __stcg(fr + j, 0);
  }
  __syncwarp();
  int32_t r = m;
  auto run_start = [&](int32_t j) -> int32_t { return __ldcg(cur + j); };
  auto run_end = [&](int32_t j) -> int32_t { return j + 1 < r ? __ldcg(cur + j + 1) : m; };
  int32_t r0 = -1;  // number of flat clusters once the tolerance-bounded rounds are over
  while (r > 1) {
    const bool bounded = r0 < 0;
    int32_t r_new = 0;
    bool merged_any = false;
    for (int32_t base = 0; base < r; base += 32) {
      const int32_t j = base + lane;
      const bool valid = j < r;
      bool merge = false;
      if (valid && j >= 1) {
        const int32_t sa = run_start(j - 1), sb = run_start(j), eb = run_end(j);
        const double d = split_distance(v[sa], v[eb - 1], tol_mode);
        if (!bounded || d <= tol) {
          merge = true;
          if (j >= 2) {
            const double dl = split_distance(v[run_start(j - 2)], v[sb - 1], tol_mode);
            if (!(d < dl)) merge = false;
          }
          if (merge && j + 1 < r) {
            const double dr = split_distance(v[sb], v[run_end(j + 1) - 1], tol_mode);
            if (!(d <= dr)) merge = false;
          }
        }
        if (merge && !bounded) {
          const int32_t fa = __ldcg(fr + sa), fb = __ldcg(fr + sb);
          const int32_t fe = eb < m ? __ldcg(fr + eb) : r0;
          const int32_t ca = fb - fa, cb = fe - fb;
          const bool single_a = sb - sa == 1, single_b = eb - sb == 1;
          const int32_t off_a = single_a ? (single_b ? 0 : cb) : 0;
          const int32_t off_b = single_b ? (single_a ? 1 : ca) : (single_a ? 0 : ca);
          if (off_a) {
            atomicAdd(dl_ + fa, off_a);
            atomicAdd(dl_ + fb, -off_a);
          }
          if (off_b) {
            atomicAdd(dl_ + fb, off_b);
            atomicAdd(dl_ + fe, -off_b);
          }
        }
      }
      const bool keep = valid && !merge;
      const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
      if (keep) __stcg(nxt + r_new + __popc(ballot & lt), run_start(j));
      r_new += __popc(ballot);
      merged_any |= __any_sync(0xffffffffu, merge);
    }
    __syncwarp();
    if (merged_any) {
      int32_t* t = cur; cur = nxt; nxt = t;
      r = r_new;
    }
    if (bounded && (!merged_any || r == 1)) {
      // flat clusters fixed: rank of every element's flat cluster, cleared difference array
      r0 = r;
      for (int32_t j = lane; j < r; j += 32) __stcg(fr + __ldcg(cur + j), 1);
      for (int32_t j = lane; j <= r; j += 32) __stcg(dl_ + j, 0);
      __syncwarp();
      int32_t carry = -1;
      for (int32_t base = 0; base < m; base += 32) {
        const int32_t j = base + lane;
        const bool h = j < m && __ldcg(fr + j) != 0;
        const uint32_t hb = __ballot_sync(0xffffffffu, h);
        if (j < m) __stcg(fr + j, carry + __popc(hb & lt) + (h ? 1 : 0));
        carry += __popc(hb);
      }
      __syncwarp();
    }
  }
  // ids of the flat clusters = inclusive prefix sum of the difference array
  int32_t carry = 0;
  for (int32_t base = 0; base < r0; base += 32) {
    const int32_t j = base + lane;
    int32_t x = j < r0 ? __ldcg(dl_ + j) : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    x += carry;
    if (j < r0) __stcg(dl_ + j, x);
    carry = __shfl_sync(0xffffffffu, x, 31);
  }
  __syncwarp();
  for (int32_t j = lane; j < m; j += 32) aid[s + j] = __ldcg(dl_ + __ldcg(fr + j));
}

__global__ void __launch_bounds__(256)
linkage_ids_kernel(const int64_t* __restrict__ gstart, const int64_t* __restrict__ n_groups_ptr,
                   const uint32_t* __restrict__ key_sorted, const double* __restrict__ vs, int64_t n, double tol,
                   int tol_mode, int32_t* list_a, int32_t* list_b, int32_t* frank, int32_t* delta,
                   int32_t* __restrict__ aid) {
  const int64_t n_groups = *n_groups_ptr;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  for (int64_t g = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps_total) {
    linkage_ids_one(g, n_groups, gstart, key_sorted, vs, n, tol, tol_mode, list_a, list_b, frank, delta, aid);
    __syncwarp();
  }
}

// Sorted position i of the RT order: combined key (label, a_mz * 2 + a_rt * 3).
__global__ void split_combine_kernel(const uint32_t* __restrict__ key_sorted, const int32_t* __restrict__ perm,
                                     const int32_t* __restrict__ a_mz /*by row*/, const int32_t* __restrict__ a_rt /*by position*/,
                                     int64_t n, uint64_t* __restrict__ key64) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = key_sorted[i];
  const uint64_t c = k == kNoiseKey ? 0ull
                                     : static_cast<uint64_t>(static_cast<uint32_t>(a_mz[perm[i]])) * 2ull +
                                           static_cast<uint64_t>(static_cast<uint32_t>(a_rt[i])) * 3ull;
  key64[i] = (static_cast<uint64_t>(k) << 32) | c;
}

__global__ void scatter_by_perm_kernel(const int32_t* __restrict__ in, const int32_t* __restrict__ perm, int64_t n,
                                       int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[perm[i]] = in[i];
}

// Runs of equal combined key are the final sub-clusters (noise rows: runs of their own).
__global__ void split_runhead64_kernel(const uint64_t* __restrict__ key64_sorted, int64_t n,
                                       uint8_t* __restrict__ runhead, uint32_t* __restrict__ key32) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = key64_sorted[i];
  const uint32_t label = static_cast<uint32_t>(k >> 32);
  key32[i] = label;
  runhead[i] = (i == 0 || k != key64_sorted[i - 1] || label == kNoiseKey) ? 1 : 0;
}

struct SplitLayout {
  uint32_t* key_a;
  uint32_t* key_b;
  int32_t* idx_a;
  int32_t* idx_b;
  uint64_t* mzkey_a;
  uint64_t* mzkey_b;
  double* vs;
  uint8_t* ghead;
  uint8_t* runhead;
  int64_t* gstart;
  int64_t* rstart;
  int64_t* n_groups;
  int64_t* n_runs;
  int32_t* list_a;
  int32_t* list_b;
  int32_t* run_i32;
  int32_t* run_id;
  int32_t* kept;
  int32_t* new_id;
  int32_t* frank;   // RT cut only
  int32_t* delta;
  int32_t* aid;
  int32_t* a_mz;
  void* cub_tmp;
  size_t cub_bytes;
};

static void split_layout(Workspace& ws, int64_t n, bool with_rt, SplitLayout& L) {
  L.key_a = ws.take<uint32_t>(n);
  L.key_b = ws.take<uint32_t>(n);
  L.idx_a = ws.take<int32_t>(n);
  L.idx_b = ws.take<int32_t>(n);
  L.mzkey_a = ws.take<uint64_t>(n);
  L.mzkey_b = ws.take<uint64_t>(n);
  L.vs = ws.take<double>(n);
  L.ghead = ws.take<uint8_t>(n);
  L.runhead = ws.take<uint8_t>(n);
  L.gstart = ws.take<int64_t>(n + 1);
  L.rstart = ws.take<int64_t>(n + 1);
  L.n_groups = ws.take<int64_t>(1);
  L.n_runs = ws.take<int64_t>(1);
  L.list_a = ws.take<int32_t>(n);
  L.list_b = ws.take<int32_t>(n);
  L.run_i32 = ws.take<int32_t>(n);
  L.run_id = ws.take<int32_t>(n);
  L.kept = ws.take<int32_t>(n + 1);
  L.new_id = ws.take<int32_t>(n + 1);
  L.frank = L.delta = L.aid = L.a_mz = nullptr;
  if (with_rt) {
    L.frank = ws.take<int32_t>(n);
    L.delta = ws.take<int32_t>(2 * n + 2);
    L.aid = ws.take<int32_t>(n);
    L.a_mz = ws.take<int32_t>(n);
  }
  const int num = static_cast<int>(n);
  size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr,
                                  (int32_t*)nullptr, num);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr,
                                  (int32_t*)nullptr, num);
  cub::DeviceSelect::Flagged(nullptr, b3, cub::CountingInputIterator<int64_t>(0), (uint8_t*)nullptr,
                             (int64_t*)nullptr, (int64_t*)nullptr, num);
  cub::DeviceScan::InclusiveSum(nullptr, b4, cub::TransformInputIterator<int32_t, U8ToI32, const uint8_t*>(nullptr, U8ToI32()),
                                (int32_t*)nullptr, num);
  cub::DeviceScan::ExclusiveSum(nullptr, b5, (int32_t*)nullptr, (int32_t*)nullptr, num + 1);
  size_t b = b1;
  if (b2 > b) b = b2;
  if (b3 > b) b = b3;
  if (b4 > b) b = b4;
  if (b5 > b) b = b5;
  L.cub_bytes = b;
  L.cub_tmp = ws.take<char>(b);
}

}  // namespace flc

extern "C" {

size_t flc_dbscan_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::DbscanLayout L;
  flc::dbscan_layout(ws, n, L);
  return ws.used + 256;
}

int flc_dbscan(const float* dist, const int32_t* indices, const int64_t* indptr, int64_t n, float eps,
               int32_t min_samples, int32_t* labels, int64_t* n_clusters, int64_t* n_clusters_dev,
               int32_t n_sweeps, int32_t* sweeps_used, int32_t* unsettled_dev,
               void* workspace, size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31) - 1, "n out of range");
  FLC_REQUIRE(min_samples >= 1, "min_samples must be >= 1");
  FLC_REQUIRE(n_clusters != nullptr || n_clusters_dev != nullptr, "no output for n_clusters");
  FLC_REQUIRE(n_sweeps >= 0 && (n_sweeps == 0 || unsettled_dev != nullptr),
              "a fixed number of sweeps needs unsettled_dev to report whether it was enough");
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    if (n_clusters) *n_clusters = 0;
    if (sweeps_used) *sweeps_used = 0;
    if (n_clusters_dev) FLC_CUDA(cudaMemsetAsync(n_clusters_dev, 0, sizeof(int64_t), stream));
    if (unsettled_dev) FLC_CUDA(cudaMemsetAsync(unsettled_dev, 0, sizeof(int32_t), stream));
    return FLC_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  DbscanLayout L;
  dbscan_layout(ws, n, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "dbscan workspace too small: need %zu", ws.used);
  const unsigned wblocks = static_cast<unsigned>((n + 255) / 256);
  const unsigned tblocks = static_cast<unsigned>((n + 1 + 255) / 256);
  timed("dbscan_core", stream, [&] { dbscan_core_kernel<<<wblocks, 256, 0, stream>>>(dist, indptr, n, eps, min_samples, L.m, L.core); });
  FLC_LAUNCH_CHECK();
  auto sweep = [&](int32_t* flag) -> int {
    FLC_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), stream));
    timed("dbscan_propagate", stream, [&] { dbscan_propagate_kernel<<<wblocks, 256, 0, stream>>>(dist, indices, indptr, n, eps, L.core, L.m,
                                                         flag); });
    FLC_LAUNCH_CHECK();
    return FLC_OK;
  };
  if (n_sweeps > 0) {
    // a fixed number of sweeps, nothing read back: the flag of the last one (did it still change something?)
    // is left in unsettled_dev for the caller's next synchronisation
    for (int i = 0; i < n_sweeps; ++i) FLC_TRY(sweep(unsettled_dev));
  } else {
    int32_t changed = 1;
    int sweeps = 0, launched = 0;
    while (changed) {
      // two sweeps before the first look at the flag (only the second one's changes count: the first
      // always changes something), one sweep per look afterwards
      for (int rep = 0; rep < (sweeps == 0 ? 2 : 1); ++rep, ++launched) FLC_TRY(sweep(L.changed));
      FLC_CUDA(cudaMemcpyAsync(&changed, L.changed, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
      FLC_CUDA(cudaStreamSynchronize(stream));
      if (++sweeps > 100000) return set_error(FLC_ERR_CUDA, "dbscan propagation did not converge");
    }
    if (sweeps_used) *sweeps_used = launched;
    if (unsettled_dev) FLC_CUDA(cudaMemsetAsync(unsettled_dev, 0, sizeof(int32_t), stream));
  }
  timed("dbscan_seed", stream, [&] { dbscan_seed_kernel<<<tblocks, 256, 0, stream>>>(L.m, L.core, n, L.seed); });
  FLC_LAUNCH_CHECK();
  size_t tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.seed, L.rank, static_cast<int>(n + 1), stream));
  count_launch(2);
  timed("dbscan_label", stream, [&] { dbscan_label_kernel<<<tblocks, 256, 0, stream>>>(L.m, L.rank, n, labels, n_clusters_dev); });
  FLC_LAUNCH_CHECK();
  if (n_clusters) {  // NULL: no synchronisation, the count stays on the device
    int32_t total = 0;
    FLC_CUDA(cudaMemcpyAsync(&total, L.rank + n, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    FLC_CUDA(cudaStreamSynchronize(stream));
    *n_clusters = total;
  }
  return FLC_OK;
}

size_t flc_split_workspace_bytes(int64_t n, int with_rt) {
  if (n <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::SplitLayout L;
  flc::split_layout(ws, n, with_rt != 0, L);
  return ws.used + 256;
}

int flc_split_clusters(const int32_t* labels_in, const double* precursor_mz, const double* rt, int64_t n,
                       double tol, int tol_mode, double rt_tol, int32_t min_samples, int values_sorted,
                       int32_t* labels_out, int64_t* n_clusters, int64_t* n_clusters_dev, void* workspace,
                       size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31) - 1, "n out of range");
  FLC_REQUIRE(tol_mode == FLC_TOL_DA || tol_mode == FLC_TOL_PPM, "Unknown precursor tolerance mode");
  FLC_REQUIRE(n_clusters != nullptr || n_clusters_dev != nullptr, "no output for n_clusters");
  const bool with_rt = rt_tol >= 0.0;
  FLC_REQUIRE(!with_rt || rt != nullptr, "rt_tol is set but no retention times were given");
  // combined key = label << 32 | (a_mz * 2 + a_rt * 3), a_* < n
  FLC_REQUIRE(!with_rt || n < (int64_t(1) << 29), "n out of range for the retention-time cut");
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    if (n_clusters) *n_clusters = 0;
    if (n_clusters_dev) FLC_CUDA(cudaMemsetAsync(n_clusters_dev, 0, sizeof(int64_t), stream));
    return FLC_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  SplitLayout L;
  split_layout(ws, n, with_rt, L);
  if (!ws.ok) return set_error(FLC_ERR_WORKSPACE, "split workspace too small: need %zu", ws.used);
  const int num = static_cast<int>(n);
  const unsigned tblocks = static_cast<unsigned>((n + 255) / 256);
  int label_bits = 1;
  while (label_bits < 32 && (int64_t(1) << label_bits) <= n) ++label_bits;  // labels_in[i] < n
  label_bits = std::min(32, label_bits + 1);
  size_t tmp;
  // Rows grouped by label, ascending in `values` inside a group (stable sorts): L.key_b = sorted
  // keys, *perm_out = sorted position -> row, L.vs = values, L.ghead / L.gstart / L.n_groups.
  auto group_rows = [&](const double* values, bool sorted, const int32_t** perm_out) -> int {
    const int32_t* idx_in;
    if (!sorted) {
      timed("split_mzkey", stream, [&] { split_mzkey_kernel<<<tblocks, 256, 0, stream>>>(values, n, L.mzkey_a, L.idx_a); });
      FLC_LAUNCH_CHECK();
      size_t t = L.cub_bytes;
      FLC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, t, L.mzkey_a, L.mzkey_b, L.idx_a, L.idx_b, num, 0, 64,
                                               stream));
      count_launch(9);
      // key of the value-sorted rows
      timed("split_key", stream, [&] { split_key_kernel<<<tblocks, 256, 0, stream>>>(labels_in, n, L.key_b, L.idx_a); });
      FLC_LAUNCH_CHECK();
      FLC_TRY(flc_gather(L.key_b, L.idx_b, n, 4, L.key_a, stream_));
      idx_in = L.idx_b;
    } else {
      timed("split_key", stream, [&] { split_key_kernel<<<tblocks, 256, 0, stream>>>(labels_in, n, L.key_a, L.idx_a); });
      FLC_LAUNCH_CHECK();
      idx_in = L.idx_a;
    }
    int32_t* perm = (idx_in == L.idx_a) ? L.idx_b : L.idx_a;
    size_t t = L.cub_bytes;
    // labels are < n: only their significant bits are sorted on (the noise key is all ones, so it still comes last)
    FLC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, t, L.key_a, L.key_b, idx_in, perm, num, 0, label_bits, stream));
    count_launch(5);
    timed("split_prepare", stream, [&] { split_prepare_kernel<<<tblocks, 256, 0, stream>>>(L.key_b, perm, values, n, L.vs, L.ghead,
                                                      L.runhead); });
    FLC_LAUNCH_CHECK();
    t = L.cub_bytes;
    FLC_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, t, cub::CountingInputIterator<int64_t>(0), L.ghead, L.gstart,
                                        L.n_groups, num, stream));
    count_launch(2);
    *perm_out = perm;
    return FLC_OK;
  };
  const int32_t* perm = nullptr;
  FLC_TRY(group_rows(precursor_mz, values_sorted != 0, &perm));
  const uint32_t* key_sorted = L.key_b;
  // The number of groups is only known on the device: a resident grid of warps strides over them.
  const unsigned gblocks = static_cast<unsigned>(std::min<int64_t>((n * 8 + 255) / 256 + 1, static_cast<int64_t>(kNumSMs) * 8));
  if (!with_rt) {
    timed("split_group", stream, [&] { split_group_kernel<<<gblocks, 256, 0, stream>>>(L.gstart, L.n_groups, key_sorted, L.vs, n, tol, tol_mode,
                                                    L.list_a, L.list_b, L.runhead); });
    FLC_LAUNCH_CHECK();
  } else {
    // fcluster ids of the m/z cut (by row), then of the RT cut (by position in RT order),
    // combined and sorted: equal (label, a_mz * 2 + a_rt * 3) = one sub-cluster
    timed("linkage_ids", stream, [&] { linkage_ids_kernel<<<gblocks, 256, 0, stream>>>(L.gstart, L.n_groups, key_sorted, L.vs, n, tol, tol_mode,
                                                    L.list_a, L.list_b, L.frank, L.delta, L.aid); });
    FLC_LAUNCH_CHECK();
    timed("split_scatter", stream, [&] { scatter_by_perm_kernel<<<tblocks, 256, 0, stream>>>(L.aid, perm, n, L.a_mz); });
    FLC_LAUNCH_CHECK();
    FLC_TRY(group_rows(rt, false, &perm));
    timed("linkage_ids", stream, [&] { linkage_ids_kernel<<<gblocks, 256, 0, stream>>>(L.gstart, L.n_groups, key_sorted, L.vs, n, rt_tol, FLC_TOL_DA,
                                                    L.list_a, L.list_b, L.frank, L.delta, L.aid); });
    FLC_LAUNCH_CHECK();
    timed("split_combine", stream, [&] { split_combine_kernel<<<tblocks, 256, 0, stream>>>(key_sorted, perm, L.a_mz, L.aid, n, L.mzkey_a); });
    FLC_LAUNCH_CHECK();
    int32_t* perm_f = (perm == L.idx_a) ? L.idx_b : L.idx_a;
    tmp = L.cub_bytes;
    FLC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, L.mzkey_a, L.mzkey_b, perm, perm_f, num, 0, 64, stream));
    count_launch(9);
    perm = perm_f;
    timed("split_runhead", stream, [&] { split_runhead64_kernel<<<tblocks, 256, 0, stream>>>(L.mzkey_b, n, L.runhead, L.key_a); });
    FLC_LAUNCH_CHECK();
    key_sorted = L.key_a;
  }
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, tmp, cub::CountingInputIterator<int64_t>(0), L.runhead,
                                      L.rstart, L.n_runs, num, stream));
  count_launch(2);
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::InclusiveSum(L.cub_tmp, tmp, cub::TransformInputIterator<int32_t, U8ToI32, const uint8_t*>(L.runhead, U8ToI32()),
                                         L.run_id, num, stream));
  count_launch(2);
  timed("split_keep", stream, [&] { split_keep_kernel<<<static_cast<unsigned>((n + 1 + 255) / 256), 256, 0, stream>>>(L.rstart, L.n_runs,
                                                                                   key_sorted, n, min_samples,
                                                                                   L.kept); });
  FLC_LAUNCH_CHECK();
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.kept, L.new_id, num + 1, stream));
  count_launch(2);
  timed("split_label", stream, [&] { split_label_kernel<<<tblocks, 256, 0, stream>>>(L.run_id, L.kept, L.new_id, perm, n, labels_out,
                                                  n_clusters_dev); });
  FLC_LAUNCH_CHECK();
  if (n_clusters) {  // NULL: no synchronisation, the count stays on the device
    int32_t total = 0;
    FLC_CUDA(cudaMemcpyAsync(&total, L.new_id + n, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    FLC_CUDA(cudaStreamSynchronize(stream));
    *n_clusters = total;
  }
  return FLC_OK;
}

}  // extern "C"
