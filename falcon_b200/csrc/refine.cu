// Stage 3b: candidate pairs -> exact re-score -> top-k -> precursor filter -> CSR.
//
// The tensor-core scan (scan_tc.cu) produces (query, candidate) pairs whose
// bf16 inner product clears `1 - eps - margin`.  This file groups them by
// query (histogram + scan + scatter), and with one warp per query
//   * checks inverted-list membership (candidate's list is one the query probes),
//   * recomputes the inner product exactly: float32 inputs, float64 accumulate,
//     rounded once to float32 (order independent, so it matches the oracle bit
//     for bit; faiss' float32 SIMD sum differs by ~1e-7 < the 1e-5 bar),
//   * applies the eps cut (dist = max(1 - ip, 0) <= eps),
//   * orders candidates by (ip descending, id ascending) with a warp bitonic
//     sort and keeps n_neighbors_ann (faiss index.search semantics, A.2),
//   * applies the precursor m/z / RT tolerance filter in that order and keeps
//     n_neighbors (A.2 _filter_neighbors_mz),
// then builds the CSR arrays with a prefix sum over the row counts.
//
// HBM-bound: algorithmic bytes = pairs * 8 (read) + pairs * 4 * low_dim worst
// case for the re-score rows (L2-resident in practice: a bucket's rows were just
// read by the scan) + N * 16 (m/z, RT) + nnz * 8 + (N + 1) * 8.
#include <cub/cub.cuh>

#include "common.cuh"

namespace flc {

constexpr int kRefineWarps = 8;
constexpr uint64_t kKeyMax = ~uint64_t(0);
constexpr int kWarpQueries = 4;   // refine_block_kernel: queries a warp handles at a time
constexpr int kWarpPairs = 64;    // candidate pairs of one query group it keeps in shared memory
constexpr int kBlockWarps = 8;

// qo[r] with a run-time r out of a register array (unrolled select)
__device__ __forceinline__ int qo_at(const int (&qo)[kWarpQueries + 1], int r) {
  int v = qo[0];
#pragma unroll
  for (int t = 1; t <= kWarpQueries; ++t) v = (r == t) ? qo[t] : v;
  return v;
}

// query_mask (nullable): queries with a zero byte keep no pairs -- their rows of the matrix stay empty
__global__ void pair_hist_kernel(const uint64_t* __restrict__ pairs, const uint64_t* __restrict__ pair_count,
                                 uint64_t capacity, int64_t n, const uint8_t* __restrict__ query_mask,
                                 uint32_t* __restrict__ cnt) {
  const uint64_t total = min(*pair_count, capacity);
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t q = static_cast<uint32_t>(pairs[i] >> 32);
    if (q < n && (query_mask == nullptr || query_mask[q])) atomicAdd(cnt + q, 1u);
  }
}

__global__ void pair_scatter_kernel(const uint64_t* __restrict__ pairs, const uint64_t* __restrict__ pair_count,
                                    uint64_t capacity, int64_t n, const uint8_t* __restrict__ query_mask,
                                    const int64_t* __restrict__ off,
                                    uint32_t* __restrict__ cursor, uint64_t* __restrict__ grouped) {
  const uint64_t total = min(*pair_count, capacity);
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t p = pairs[i];
    const uint32_t q = static_cast<uint32_t>(p >> 32);
    if (q < n && (query_mask == nullptr || query_mask[q])) {
      const uint32_t slot = atomicAdd(cursor + q, 1u);
      grouped[off[q] + slot] = p & 0xffffffffull;
    }
  }
}

// Order-preserving float -> uint (ascending), then inverted so that a larger
// inner product gives a smaller key.
__device__ __forceinline__ uint32_t ip_key_desc(float ip) {
  uint32_t u = __float_as_uint(ip);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}
__device__ __forceinline__ float ip_from_key(uint32_t k) {
  uint32_t u = ~k;
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), m);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Ascending bitonic sort of one key per lane.
__device__ __forceinline__ uint64_t warp_sort32(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t other = shfl_xor_u64(key, j);
      const bool up = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      const bool take_min = (up == lower);
      key = take_min ? (key < other ? key : other) : (key > other ? key : other);
    }
  }
  return key;
}

// Ascending bitonic sort of `len` (power of two >= 64) keys in shared memory by one warp.
__device__ __forceinline__ void warp_sort_smem(uint64_t* buf, int len, int lane) {
  for (int k = 2; k <= len; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (len >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int p = i | j;
        const bool up = ((i & k) == 0);
        const uint64_t a = buf[i], b = buf[p];
        if ((a > b) == up) {
          buf[i] = b;
          buf[p] = a;
        }
      }
      __syncwarp();
    }
  }
}

struct RefineParams {
  const uint16_t* ell_idx;  // nullable sparse copy of the rows
  const float* ell_val;
  int32_t ell_width;
  const float* x;
  int64_t ld;
  int64_t n;
  uint32_t low_dim;
  const double* mz;
  const float* rt;
  const int32_t* list_id;
  const int32_t* probes;
  int32_t max_nprobe;
  double tol;
  int tol_mode;
  double rt_tol;
  int32_t k;
  int32_t k_ann;
  int32_t ka_pow2;  // power of two >= max(k_ann, 32)
  float eps;
  int use_eps;
};

__device__ __forceinline__ bool tolerance_ok(const RefineParams& P, double mq, float rq, double mc, float rc) {
  const double dm = fabs(mq - mc);
  bool ok = P.tol_mode == FLC_TOL_DA ? (dm < P.tol) : (dm / mc * 1000000.0 < P.tol);
  if (ok && P.rt != nullptr && P.rt_tol >= 0.0)
    ok = fabs(static_cast<double>(rq) - static_cast<double>(rc)) < P.rt_tol;
  return ok;
}

// Exact inner product of candidate row c with the query row held densely in
// shared memory (xq): float32 inputs, float64 accumulate, rounded once.  Sparse
// rows: every lane takes two ELL slots (zero padding multiplies to an exact zero).
__device__ __forceinline__ float exact_ip(const RefineParams& P, const float* xq, uint32_t c, int lane) {
  double acc = 0.0;
  if (P.ell_idx != nullptr) {
    const int64_t rbase = static_cast<int64_t>(c) * P.ell_width;
    for (int32_t j = 2 * lane; j < P.ell_width; j += 64) {
      const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(P.ell_idx + rbase + j));
      const float2 vv = __ldg(reinterpret_cast<const float2*>(P.ell_val + rbase + j));
      acc = fma(static_cast<double>(vv.x), static_cast<double>(xq[kk & 0xffffu]), acc);
      acc = fma(static_cast<double>(vv.y), static_cast<double>(xq[kk >> 16]), acc);
    }
  } else {
    const float* xc = P.x + static_cast<int64_t>(c) * P.ld;
    for (uint32_t i = lane; i < P.low_dim; i += 32)
      acc = fma(static_cast<double>(xq[i]), static_cast<double>(__ldg(xc + i)), acc);
  }
  return static_cast<float>(warp_sum_f64(acc));
}

// One warp per query (grid-stride).  The query row sits densely in shared memory:
// scattered in from its sparse copy and scattered back out to zero afterwards.
__global__ void __launch_bounds__(kRefineWarps * 32)
refine_kernel(RefineParams P, const int64_t* __restrict__ off, uint64_t* __restrict__ grouped,
              int32_t* __restrict__ row_count, const int32_t* __restrict__ deferred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t below = (1u << lane) - 1u;
  const size_t per_warp = static_cast<size_t>(2) * P.ka_pow2 * sizeof(uint64_t) +
                          ((static_cast<size_t>(P.low_dim) * sizeof(float) + 15) & ~size_t(15));
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw + warp * per_warp);
  float* xq = reinterpret_cast<float*>(buf + 2 * P.ka_pow2);
  const bool sparse = P.ell_idx != nullptr;
  // With `deferred` this kernel only finishes the query groups refine_block_kernel passed on:
  // deferred[0] = how many, deferred[1 ..] = their group numbers (kWarpQueries queries each).
  const int64_t n_work = deferred != nullptr ? static_cast<int64_t>(deferred[0]) * kWarpQueries : P.n;
  if (n_work == 0) return;
  if (sparse) {
    for (uint32_t i = lane; i < P.low_dim; i += 32) xq[i] = 0.f;
    __syncwarp();
  }
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kRefineWarps;
  for (int64_t w = static_cast<int64_t>(blockIdx.x) * kRefineWarps + warp; w < n_work; w += warps_total) {
    const int64_t q = deferred != nullptr
                          ? static_cast<int64_t>(deferred[1 + w / kWarpQueries]) * kWarpQueries + (w % kWarpQueries)
                          : w;
    if (q >= P.n) continue;
    const int64_t base = off[q];
    const int64_t m = off[q + 1] - base;
    if (m == 0) {
      if (lane == 0) row_count[q] = 0;
      continue;
    }
    // ---- query row -> dense
    if (sparse) {
      const int64_t qb = q * P.ell_width;
      for (int32_t j = 2 * lane; j < P.ell_width; j += 64) {
        const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(P.ell_idx + qb + j));
        const float2 vv = __ldg(reinterpret_cast<const float2*>(P.ell_val + qb + j));
        if (vv.x != 0.f) xq[kk & 0xffffu] = vv.x;
        if (vv.y != 0.f) xq[kk >> 16] = vv.y;
      }
    } else {
      const float* xrow = P.x + q * P.ld;
      for (uint32_t i = lane; i < P.low_dim; i += 32) xq[i] = xrow[i];
    }
    __syncwarp();
    const double mq = P.mz[q];
    const float rq = P.rt ? P.rt[q] : 0.f;
    int32_t kept = 0;

    if (m <= 32) {
      // ---- every candidate in its own lane: ids, membership, tolerance operands up front
      const bool have = lane < m;
      const uint32_t c = have ? static_cast<uint32_t>(grouped[base + lane]) : 0u;
      bool member = have;
      if (have && P.list_id != nullptr) {
        const int32_t lc = P.list_id[c];
        bool hit = false;
        for (int32_t t = 0; t < P.max_nprobe; ++t) hit |= (__ldg(P.probes + q * P.max_nprobe + t) == lc);
        member = hit;
      }
      const double mc = have ? P.mz[c] : 1.0;
      const float rc = (have && P.rt) ? P.rt[c] : 0.f;
      const uint32_t members = __ballot_sync(0xffffffffu, member);
      float ip = 0.f;
      for (uint32_t rest = members; rest != 0u; rest &= rest - 1u) {
        const int j = __ffs(rest) - 1;
        const float v = exact_ip(P, xq, __shfl_sync(0xffffffffu, c, j), lane);
        if (lane == j) ip = v;
      }
      const float dist = fmaxf(1.0f - ip, 0.0f);
      const bool valid = member && (!P.use_eps || dist <= P.eps);
      const uint64_t key = valid ? ((static_cast<uint64_t>(ip_key_desc(ip)) << 32) | c) : kKeyMax;
      const bool ok = valid && tolerance_ok(P, mq, rq, mc, rc);
      // ---- rank by counting: lanes with a smaller (ip desc, id asc) key
      uint32_t lt = 0;
      for (int i = 0; i < m; ++i) {
        const uint64_t ki = (static_cast<uint64_t>(__shfl_sync(0xffffffffu, static_cast<uint32_t>(key >> 32), i)) << 32) |
                            __shfl_sync(0xffffffffu, static_cast<uint32_t>(key), i);
        lt |= (ki < key ? 1u : 0u) << i;
      }
      const uint32_t ann = __ballot_sync(0xffffffffu, valid && __popc(lt) < P.k_ann);
      const uint32_t okm = __ballot_sync(0xffffffffu, ok) & ann;
      const int prank = __popc(lt & okm);
      if (ok && ((ann >> lane) & 1u) && prank < P.k)
        grouped[base + prank] = (static_cast<uint64_t>(__float_as_uint(dist)) << 32) | c;
      kept = min(__popc(okm), P.k);
    } else {
      // ---- Phase A: exact score of every candidate -> sort key (in place)
      for (int64_t j0 = 0; j0 < m; j0 += 32) {
        const bool have = j0 + lane < m;
        const uint32_t c = have ? static_cast<uint32_t>(grouped[base + j0 + lane]) : 0u;
        bool member = have;
        if (have && P.list_id != nullptr) {
          const int32_t lc = P.list_id[c];
          bool hit = false;
          for (int32_t t = 0; t < P.max_nprobe; ++t) hit |= (__ldg(P.probes + q * P.max_nprobe + t) == lc);
          member = hit;
        }
        float ip = 0.f;
        for (uint32_t rest = __ballot_sync(0xffffffffu, member); rest != 0u; rest &= rest - 1u) {
          const int j = __ffs(rest) - 1;
          const float v = exact_ip(P, xq, __shfl_sync(0xffffffffu, c, j), lane);
          if (lane == j) ip = v;
        }
        const float dist = fmaxf(1.0f - ip, 0.0f);
        const bool valid = member && (!P.use_eps || dist <= P.eps);
        if (have) grouped[base + j0 + lane] = valid ? ((static_cast<uint64_t>(ip_key_desc(ip)) << 32) | c) : kKeyMax;
      }
      __syncwarp();
      // ---- Phase B: running top-k_ann by bitonic merges in shared memory
      const int KA = P.ka_pow2;
      for (int t = lane; t < KA; t += 32) buf[t] = kKeyMax;
      for (int64_t blk = 0; blk < m; blk += KA) {
        for (int t = lane; t < KA; t += 32) buf[KA + t] = (blk + t < m) ? grouped[base + blk + t] : kKeyMax;
        __syncwarp();
        warp_sort_smem(buf, 2 * KA, lane);
      }
      // ---- Phase C: tolerance filter in similarity order
      for (int s = 0; s < P.k_ann && kept < P.k; s += 32) {
        const int t = s + lane;
        const uint64_t key = (t < P.k_ann && t < KA) ? buf[t] : kKeyMax;
        const bool valid = key != kKeyMax;
        const uint32_t c = static_cast<uint32_t>(key);
        const bool pass = valid && tolerance_ok(P, mq, rq, P.mz[c], P.rt ? P.rt[c] : 0.f);
        const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
        const int rank = kept + __popc(ballot & below);
        if (pass && rank < P.k) {
          const float ip = ip_from_key(static_cast<uint32_t>(key >> 32));
          const float dist = fmaxf(1.0f - ip, 0.0f);
          grouped[base + rank] = (static_cast<uint64_t>(__float_as_uint(dist)) << 32) | c;
        }
        kept = min(kept + __popc(ballot), P.k);
        if (__ballot_sync(0xffffffffu, valid) != 0xffffffffu) break;
      }
    }
    if (lane == 0) row_count[q] = kept;
    // ---- query row back to zero
    if (sparse) {
      __syncwarp();
      const int64_t qb = q * P.ell_width;
      for (int32_t j = 2 * lane; j < P.ell_width; j += 64) {
        const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(P.ell_idx + qb + j));
        xq[kk & 0xffffu] = 0.f;
        xq[kk >> 16] = 0.f;
      }
    }
    __syncwarp();
  }
}

// Fast path for sparse rows: every WARP takes four consecutive queries at a time (no
// block-wide barriers), holds their rows densely in its slice of shared memory
// and gives every candidate pair four lanes -- the sparse candidate row streams
// through registers (independent 16-byte loads, zero padding multiplies to an exact
// zero), products accumulate in float64, two shuffles join the quad.  Ranks come
// from counting smaller keys among the query's pairs, then the same top-k_ann /
// tolerance / first-k selection as refine_kernel.  Query groups with more than
// kWarpPairs pairs are appended to `deferred` (deferred[0] = how many) and left to
// refine_kernel.
__global__ void __launch_bounds__(kBlockWarps * 32)
refine_block_kernel(RefineParams P, const int64_t* __restrict__ off, uint64_t* __restrict__ grouped,
                    int32_t* __restrict__ row_count, int32_t* __restrict__ deferred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3;
  const size_t per_warp = static_cast<size_t>(kWarpQueries) * P.low_dim * sizeof(float) + kWarpPairs * 9 + 16;
  float* xq = reinterpret_cast<float*>(smem_raw + warp * ((per_warp + 15) & ~size_t(15)));  // [4][low_dim]
  uint64_t* keys = reinterpret_cast<uint64_t*>(xq + kWarpQueries * P.low_dim);           // [kWarpPairs]
  uint8_t* flags = reinterpret_cast<uint8_t*>(keys + kWarpPairs);  // bit 0: passes tolerance, bit 1: and within k_ann
  const int W = P.ell_width;
  const int cpr = W >> 3;  // 8-slot chunks per row
  const int64_t n_groups = (P.n + kWarpQueries - 1) / kWarpQueries;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kBlockWarps;
  for (uint32_t i = lane; i < kWarpQueries * P.low_dim; i += 32) xq[i] = 0.f;
  __syncwarp();

  for (int64_t grp = static_cast<int64_t>(blockIdx.x) * kBlockWarps + warp; grp < n_groups; grp += warps_total) {
    const int64_t q0 = grp * kWarpQueries;
    const int nq = static_cast<int>(min(static_cast<int64_t>(kWarpQueries), P.n - q0));
    // offsets of the group's queries: lane r holds off[q0 + r] - off[q0]
    const int64_t my_off = off[q0 + min(lane, nq)];
    const int64_t base = __shfl_sync(0xffffffffu, my_off, 0);
    const int64_t pb64 = __shfl_sync(0xffffffffu, my_off, nq) - base;
    if (pb64 > kWarpPairs) {  // uniform
      if (lane == 0) deferred[1 + atomicAdd(deferred, 1)] = static_cast<int32_t>(grp);
      continue;
    }
    const int pb = static_cast<int>(pb64);
    const int rel = static_cast<int>(my_off - base);
    int qo[kWarpQueries + 1];
#pragma unroll
    for (int r = 0; r <= kWarpQueries; ++r) qo[r] = __shfl_sync(0xffffffffu, rel, min(r, nq));
    if (pb == 0) {
      if (lane < nq) row_count[q0 + lane] = 0;
      continue;
    }
    auto query_of = [&](int p) {  // last r with qo[r] <= p
      int r = 0;
#pragma unroll
      for (int t = 1; t < kWarpQueries; ++t) r += (t < nq && qo[t] <= p) ? 1 : 0;
      return r;
    };
    // ---- query rows -> dense
    for (int it = lane; it < nq * cpr; it += 32) {
      const int r = it / cpr, j0 = (it - r * cpr) << 3;
      const int64_t g = (q0 + r) * W + j0;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(P.ell_val + g));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(P.ell_val + g + 4));
      const uint4 ki = __ldg(reinterpret_cast<const uint4*>(P.ell_idx + g));
      float* xr = xq + r * P.low_dim;
      const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const uint32_t kk[4] = {ki.x, ki.y, ki.z, ki.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (vv[u] != 0.f) xr[(kk[u >> 1] >> ((u & 1) * 16)) & 0xffffu] = vv[u];
    }
    __syncwarp();
    // ---- four lanes per candidate pair: membership, exact score, tolerance
    for (int p0 = 0; p0 < pb; p0 += 8) {
      const int p = p0 + (lane >> 2);
      const bool have = p < pb;
      int lo = 0;
      uint32_t c = 0;
      double acc = 0.0;
      if (have) {
        lo = query_of(p);
        c = static_cast<uint32_t>(grouped[base + p]);
        const float* xr = xq + lo * P.low_dim;
        const int64_t rb = static_cast<int64_t>(c) * W;
#pragma unroll 2
        for (int j0 = 8 * sub; j0 < W; j0 += 32) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(P.ell_val + rb + j0));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(P.ell_val + rb + j0 + 4));
          const uint4 ki = __ldg(reinterpret_cast<const uint4*>(P.ell_idx + rb + j0));
          acc = fma(static_cast<double>(v0.x), static_cast<double>(xr[ki.x & 0xffffu]), acc);
          acc = fma(static_cast<double>(v0.y), static_cast<double>(xr[ki.x >> 16]), acc);
          acc = fma(static_cast<double>(v0.z), static_cast<double>(xr[ki.y & 0xffffu]), acc);
          acc = fma(static_cast<double>(v0.w), static_cast<double>(xr[ki.y >> 16]), acc);
          acc = fma(static_cast<double>(v1.x), static_cast<double>(xr[ki.z & 0xffffu]), acc);
          acc = fma(static_cast<double>(v1.y), static_cast<double>(xr[ki.z >> 16]), acc);
          acc = fma(static_cast<double>(v1.z), static_cast<double>(xr[ki.w & 0xffffu]), acc);
          acc = fma(static_cast<double>(v1.w), static_cast<double>(xr[ki.w >> 16]), acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (have && sub == 0) {
        const int64_t q = q0 + lo;
        bool member = true;
        if (P.list_id != nullptr) {
          const int32_t lc = __ldg(P.list_id + c);
          bool hit = false;
          for (int32_t t = 0; t < P.max_nprobe; ++t) hit |= (__ldg(P.probes + q * P.max_nprobe + t) == lc);
          member = hit;
        }
        uint64_t key = kKeyMax;
        uint8_t flag = 0;
        const float ip = static_cast<float>(acc);
        const float dist = fmaxf(1.0f - ip, 0.0f);
        if (member && (!P.use_eps || dist <= P.eps)) {
          key = (static_cast<uint64_t>(ip_key_desc(ip)) << 32) | c;
          flag = tolerance_ok(P, P.mz[q], P.rt ? P.rt[q] : 0.f, P.mz[c], P.rt ? P.rt[c] : 0.f) ? 1 : 0;
        }
        keys[p] = key;
        flags[p] = flag;
      }
    }
    __syncwarp();
    // ---- rank among the query's pairs; only the k_ann best are eligible
    for (int p = lane; p < pb; p += 32) {
      const uint64_t key = keys[p];
      if (key == kKeyMax || !(flags[p] & 1)) continue;
      const int lo = query_of(p);
      int rank = 0;
      for (int j = qo_at(qo, lo); j < qo_at(qo, lo + 1); ++j) rank += keys[j] < key ? 1 : 0;
      if (rank < P.k_ann) flags[p] |= 2;
    }
    __syncwarp();
    // ---- position among the eligible ones that pass the tolerance: the first k are kept
    for (int p = lane; p < pb; p += 32) {
      if (!(flags[p] & 2)) continue;
      const uint64_t key = keys[p];
      const int lo = query_of(p);
      const int j0 = qo_at(qo, lo), j1 = qo_at(qo, lo + 1);
      int prank = 0;
      for (int j = j0; j < j1; ++j) prank += ((flags[j] & 2) && keys[j] < key) ? 1 : 0;
      if (prank < P.k) {
        const float ip = ip_from_key(static_cast<uint32_t>(key >> 32));
        const float dist = fmaxf(1.0f - ip, 0.0f);
        grouped[base + j0 + prank] = (static_cast<uint64_t>(__float_as_uint(dist)) << 32) | (key & 0xffffffffull);
      }
    }
    if (lane < nq) {
      int kept = 0;
      for (int j = qo_at(qo, lane); j < qo_at(qo, lane + 1); ++j) kept += (flags[j] & 2) ? 1 : 0;
      row_count[q0 + lane] = min(kept, P.k);
    }
    // ---- query rows back to zero
    for (int it = lane; it < nq * cpr; it += 32) {
      const int r = it / cpr, j0 = (it - r * cpr) << 3;
      const int64_t g = (q0 + r) * W + j0;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(P.ell_val + g));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(P.ell_val + g + 4));
      const uint4 ki = __ldg(reinterpret_cast<const uint4*>(P.ell_idx + g));
      float* xr = xq + r * P.low_dim;
      const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const uint32_t kk[4] = {ki.x, ki.y, ki.z, ki.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)  // padding slots carry column 0: only real entries are cleared
        if (vv[u] != 0.f) xr[(kk[u >> 1] >> ((u & 1) * 16)) & 0xffffu] = 0.f;
    }
    __syncwarp();
  }
}

__global__ void csr_compact_kernel(const uint64_t* __restrict__ grouped, const int64_t* __restrict__ off,
                                   const int64_t* __restrict__ indptr, int64_t n, uint64_t nnz_capacity,
                                   float* __restrict__ dist, int32_t* __restrict__ indices) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int64_t src = off[q];
  const int64_t dst = indptr[q];
  const int64_t cnt = indptr[q + 1] - dst;
  for (int64_t t = 0; t < cnt; ++t) {
    if (static_cast<uint64_t>(dst + t) < nnz_capacity) {
      const uint64_t e = grouped[src + t];
      dist[dst + t] = __uint_as_float(static_cast<uint32_t>(e >> 32));
      indices[dst + t] = static_cast<int32_t>(static_cast<uint32_t>(e));
    }
  }
}

struct KnnLayout {
  uint32_t* cnt;     // [n + 1]
  uint32_t* cursor;  // [n]
  int64_t* off;      // [n + 1]
  int32_t* row_count;  // [n + 1]
  uint64_t* grouped;   // [n_pairs]
  int32_t* deferred;   // [1 + query blocks]
  void* cub_tmp;
  size_t cub_bytes;
};

static void knn_layout(Workspace& ws, int64_t n, uint64_t n_pairs, KnnLayout& L) {
  L.cnt = ws.take<uint32_t>(n + 1);
  L.cursor = ws.take<uint32_t>(n + 1);
  L.off = ws.take<int64_t>(n + 1);
  L.row_count = ws.take<int32_t>(n + 1);
  L.grouped = ws.take<uint64_t>(n_pairs ? n_pairs : 1);
  L.deferred = ws.take<int32_t>(2 + static_cast<size_t>((n + kWarpQueries - 1) / kWarpQueries));
  size_t b1 = 0, b2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b1, (uint32_t*)nullptr, (int64_t*)nullptr, static_cast<int>(n + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, b2, (int32_t*)nullptr, (int64_t*)nullptr, static_cast<int>(n + 1));
  L.cub_bytes = b1 > b2 ? b1 : b2;
  L.cub_tmp = ws.take<char>(L.cub_bytes);
}

}  // namespace flc

extern "C" {

size_t flc_knn_csr_workspace_bytes(int64_t n, uint64_t n_pairs) {
  if (n <= 0) return 256;
  flc::Workspace ws(nullptr, 0);
  flc::KnnLayout L;
  flc::knn_layout(ws, n, n_pairs, L);
  return ws.used + 256;
}

int flc_knn_csr(const uint64_t* pairs, const uint64_t* pair_count, uint64_t pair_capacity,
                const float* x, int64_t ld, const uint16_t* ell_idx, const float* ell_val,
                int32_t ell_width, int64_t n, uint32_t low_dim,
                const double* precursor_mz, const float* rt, const int32_t* list_id,
                const int32_t* probes, int32_t max_nprobe, double tol, int tol_mode, double rt_tol,
                int32_t n_neighbors, int32_t n_neighbors_ann, float eps_cut, const uint8_t* query_mask,
                float* dist, int32_t* indices, uint64_t nnz_capacity, int64_t* indptr, int64_t* nnz,
                void* workspace, size_t workspace_bytes, flc_stream_t stream_) {
  using namespace flc;
  FLC_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "n out of range");
  FLC_REQUIRE(tol_mode == FLC_TOL_DA || tol_mode == FLC_TOL_PPM, "Unknown precursor tolerance mode");
  FLC_REQUIRE(n_neighbors > 0 && n_neighbors_ann >= n_neighbors,
              "n_neighbors_ann must be >= n_neighbors > 0");
  FLC_REQUIRE(n_neighbors_ann <= 1024, "n_neighbors_ann > 1024 not supported");
  FLC_REQUIRE((list_id == nullptr) == (probes == nullptr), "list_id and probes go together");
  FLC_REQUIRE((ell_idx == nullptr) == (ell_val == nullptr), "ell_idx and ell_val go together");
  FLC_REQUIRE(x != nullptr || ell_idx != nullptr, "need dense or ELL rows");
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    if (nnz) *nnz = 0;
    FLC_CUDA(cudaMemsetAsync(indptr, 0, sizeof(int64_t), stream));
    if (nnz) FLC_CUDA(cudaStreamSynchronize(stream));
    return FLC_OK;
  }
  // The workspace is laid out for pair_capacity pairs; whether the scan overflowed it is
  // checked with the one synchronisation at the end (this op returns nnz anyway).
  Workspace ws(workspace, workspace_bytes);
  KnnLayout L;
  knn_layout(ws, n, pair_capacity, L);
  if (!ws.ok)
    return set_error(FLC_ERR_WORKSPACE, "knn_csr workspace too small: need %zu bytes for %llu pairs",
                     ws.used, static_cast<unsigned long long>(pair_capacity));
  const uint64_t total = pair_capacity;  // upper bound for grid sizing; kernels read *pair_count
  FLC_CUDA(cudaMemsetAsync(L.cnt, 0, sizeof(uint32_t) * (n + 1), stream));
  FLC_CUDA(cudaMemsetAsync(L.cursor, 0, sizeof(uint32_t) * (n + 1), stream));
  FLC_CUDA(cudaMemsetAsync(L.row_count, 0, sizeof(int32_t) * (n + 1), stream));
  const unsigned pair_blocks =
      static_cast<unsigned>(std::min<uint64_t>((total + 255) / 256 + 1, uint64_t(kNumSMs) * 32));
  timed("pair_hist", stream, [&] { pair_hist_kernel<<<pair_blocks, 256, 0, stream>>>(pairs, pair_count, pair_capacity, n, query_mask, L.cnt); });
  FLC_LAUNCH_CHECK();
  size_t tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.cnt, L.off, static_cast<int>(n + 1), stream));
  count_launch(2);
  timed("pair_scatter", stream, [&] { pair_scatter_kernel<<<pair_blocks, 256, 0, stream>>>(pairs, pair_count, pair_capacity, n, query_mask, L.off,
                                                       L.cursor, L.grouped); });
  FLC_LAUNCH_CHECK();

  RefineParams P;
  P.ell_idx = ell_idx; P.ell_val = ell_val; P.ell_width = ell_width;
  P.x = x; P.ld = ld; P.n = n; P.low_dim = low_dim; P.mz = precursor_mz; P.rt = rt;
  P.list_id = list_id; P.probes = probes; P.max_nprobe = max_nprobe;
  P.tol = tol; P.tol_mode = tol_mode; P.rt_tol = rt_tol;
  P.k = n_neighbors; P.k_ann = n_neighbors_ann;
  int ka = 32;
  while (ka < n_neighbors_ann) ka <<= 1;
  P.ka_pow2 = ka;
  P.use_eps = !(eps_cut != eps_cut);  // NaN disables the cut
  P.eps = eps_cut;
  const size_t per_warp = static_cast<size_t>(2) * ka * sizeof(uint64_t) +
                          ((static_cast<size_t>(low_dim) * sizeof(float) + 15) & ~size_t(15));
  const size_t smem = per_warp * kRefineWarps;
  FLC_REQUIRE(smem <= 200 * 1024, "low_dim / n_neighbors_ann too large for the refine kernel");
  if (smem > 48 * 1024)
    FLC_CUDA(cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  FLC_REQUIRE(ell_idx == nullptr || ((ell_width % 2) == 0 && (reinterpret_cast<uintptr_t>(ell_idx) % 4) == 0 &&
                                     (reinterpret_cast<uintptr_t>(ell_val) % 8) == 0),
              "ELL arrays must be even-width and 4/8-byte aligned");
  const unsigned rblocks = static_cast<unsigned>(
      std::min<int64_t>((n + kRefineWarps - 1) / kRefineWarps, static_cast<int64_t>(kNumSMs) * 16));
  const int32_t* deferred = nullptr;
  const size_t warp_bytes = ((static_cast<size_t>(kWarpQueries) * low_dim * sizeof(float) + kWarpPairs * 9 + 16) + 15) & ~size_t(15);
  const size_t bsmem = warp_bytes * kBlockWarps;
  const bool block_path = ell_idx != nullptr && (ell_width % 8) == 0 && bsmem <= 100 * 1024 &&
                          (reinterpret_cast<uintptr_t>(ell_idx) % 16) == 0 && (reinterpret_cast<uintptr_t>(ell_val) % 16) == 0;
  if (block_path) {
    const int64_t n_groups = (n + kWarpQueries - 1) / kWarpQueries;
    FLC_CUDA(cudaMemsetAsync(L.deferred, 0, sizeof(int32_t), stream));  // [0] = count, [1 ..] = deferred groups
    FLC_CUDA(cudaFuncSetAttribute(refine_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(bsmem)));
    const int64_t per_sm = std::max<int64_t>(1, (227 * 1024) / static_cast<int64_t>(bsmem + 1024));
    const unsigned bblocks = static_cast<unsigned>(
        std::min<int64_t>((n_groups + kBlockWarps - 1) / kBlockWarps, static_cast<int64_t>(kNumSMs) * per_sm));
    timed("refine_block", stream, [&] { refine_block_kernel<<<bblocks, kBlockWarps * 32, bsmem, stream>>>(
        P, L.off, L.grouped, L.row_count, L.deferred); });
    FLC_LAUNCH_CHECK();
    deferred = L.deferred;
  }
  timed("refine", stream, [&] { refine_kernel<<<rblocks, kRefineWarps * 32, smem, stream>>>(
      P, L.off, L.grouped, L.row_count, deferred); });
  FLC_LAUNCH_CHECK();
  tmp = L.cub_bytes;
  FLC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.row_count, indptr, static_cast<int>(n + 1), stream));
  count_launch(2);
  if (nnz != nullptr) {
    int64_t total_nnz = 0;
    uint64_t n_pairs = 0;
    FLC_CUDA(cudaMemcpyAsync(&total_nnz, indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    FLC_CUDA(cudaMemcpyAsync(&n_pairs, pair_count, sizeof(n_pairs), cudaMemcpyDeviceToHost, stream));
    FLC_CUDA(cudaStreamSynchronize(stream));
    if (n_pairs > pair_capacity)
      return set_error(FLC_ERR_CAPACITY, "scan produced %llu candidate pairs, capacity %llu",
                       static_cast<unsigned long long>(n_pairs), static_cast<unsigned long long>(pair_capacity));
    *nnz = total_nnz;
    if (static_cast<uint64_t>(total_nnz) > nnz_capacity)
      return set_error(FLC_ERR_CAPACITY, "CSR needs %lld entries, capacity %llu",
                       static_cast<long long>(total_nnz), static_cast<unsigned long long>(nnz_capacity));
  } else {
    // No synchronisation: the caller checks *pair_count <= pair_capacity (and reads indptr[n]) when it next
    // synchronises.  nnz <= min(pairs, n * n_neighbors), so the CSR arrays cannot overflow unless the pairs did.
    FLC_REQUIRE(nnz_capacity >= std::min<uint64_t>(pair_capacity, static_cast<uint64_t>(n) * n_neighbors),
                "nnz_capacity must be at least min(pair_capacity, n * n_neighbors) without a synchronisation");
  }
  const unsigned cblocks = static_cast<unsigned>((n + 255) / 256);
  timed("csr_compact", stream, [&] { csr_compact_kernel<<<cblocks, 256, 0, stream>>>(L.grouped, L.off, indptr, n, nnz_capacity, dist, indices); });
  FLC_LAUNCH_CHECK();
  return FLC_OK;
}

}  // extern "C"
