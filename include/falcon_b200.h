/*
 * falcon_b200 -- C ABI of the B200-native falcon clustering hot path.
 *
 * Every entry point is plain C: raw device pointers + sizes + a CUDA stream, no
 * torch types.  All functions return 0 (FLC_OK) or a negative error code and
 * leave a message retrievable with flc_last_error() (thread local).  Unless
 * stated otherwise pointers are DEVICE pointers owned by the caller; the
 * library never allocates device memory behind the caller's back -- each op
 * has a *_workspace_bytes() query and takes the scratch buffer explicitly.
 * Ops are asynchronous on `stream` unless they return a host value (those
 * synchronise the stream; said in the comment).
 *
 * Reference interfaces each entry point replaces are cited as
 * /root/reference/<file>:<line>; "A.n" refers to SURVEY.md Appendix A (the
 * published falcon 0.1.x pipeline that north_star names, absent from the
 * mounted snapshot).
 */
#ifndef FALCON_B200_H_
#define FALCON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FLC_API __attribute__((visibility("default")))
#else
#define FLC_API
#endif

typedef void* flc_stream_t; /* cudaStream_t */

enum {
  FLC_OK = 0,
  FLC_ERR_INVALID = -1,   /* bad argument (reference raises ValueError) */
  FLC_ERR_CUDA = -2,      /* CUDA runtime/driver error */
  FLC_ERR_CAPACITY = -3,  /* caller-provided buffer too small; see message */
  FLC_ERR_WORKSPACE = -4, /* workspace too small */
  FLC_ERR_UNSUPPORTED = -5
};

enum { FLC_TOL_DA = 0, FLC_TOL_PPM = 1 };

/* ------------------------------------------------------------------ misc */
FLC_API const char* flc_last_error(void);
FLC_API int flc_version(void);
/* Kernels launched by this library since load / last reset (all threads). */
FLC_API uint64_t flc_launch_count(void);
FLC_API void flc_reset_launch_count(void);
/* Per-kernel device timing (CUDA events on the launching stream around every main
 * kernel).  enable(1) ... run ... count()/get(i): summed milliseconds and launches
 * per kernel name since the last reset.  count() synchronises the recorded events. */
FLC_API void flc_profile_enable(int on);
FLC_API void flc_profile_reset(void);
FLC_API int flc_profile_count(void);
FLC_API int flc_profile_get(int i, char* name, int name_bytes, double* total_ms, int* launches);
/* Compute capability of `device` must be 10.x; returns FLC_ERR_UNSUPPORTED otherwise. */
FLC_API int flc_check_device(int device);

/* ------------------------------------------------------------------ a1: get_dim
 * falcon/cluster/spectrum.py:172-199 (float32 arithmetic forced by the numba
 * signature at :172).  Host function. */
FLC_API int flc_get_dim(float min_mz, float max_mz, float bin_size,
                uint32_t* vec_len, float* start_dim, float* end_dim);

/* ------------------------------------------------------------------ a3: hash_lookup
 * A.1: hash_lookup[i] = murmurhash3_32(int32 i, seed, positive=True) % low_dim. */
FLC_API int flc_hash_table(uint32_t vec_len, uint32_t low_dim, uint32_t seed,
                   uint32_t* out /*[vec_len]*/, flc_stream_t stream);

/* ------------------------------------------------------------------ a2 + a4: vectorise
 * Binning (falcon/cluster/spectrum.py:250-296, expression :291 in float64),
 * feature hashing + L2 norm (A.1; snapshot sibling spectrum.py:202-247).
 * Row r of the outputs is spectrum order[r] (order == NULL: identity).  With
 * dest != NULL the spectrum handled as r is written to row dest[r] instead (a
 * scatter: lets a chunk of the input, in input order, land in bucket order).
 *   out_f32   [n, ld_f32]  float32 (nullable)
 *   out_bf16  [n, ld_bf16] bfloat16 bits, columns >= low_dim zeroed (nullable)
 *   out_hash_idx [n_peaks] hashed column of every peak in input peak order,
 *                -1 for peaks outside [0, vec_len) (nullable; parity checks)
 *   ell_idx / ell_val [n, ell_width] sparse (ELL) copy of the rows: non-zero
 *                columns ascending (uint16) and their float32 values, zero padded
 *                (nullable); ell_nnz [n] uint16 = populated slots of every row
 *                (nullable); *ell_overflow (device int32, caller zeroes it) receives
 *                the largest row population if one exceeds ell_width */
FLC_API int flc_vectorize(const float* mz, const float* intensity, const int64_t* indptr,
                  const int32_t* order, const int32_t* dest, int64_t n,
                  double min_mz, double bin_size, uint32_t vec_len,
                  uint32_t low_dim, uint32_t seed, int norm,
                  float* out_f32, int64_t ld_f32,
                  uint16_t* out_bf16, int64_t ld_bf16,
                  int32_t* out_hash_idx,
                  uint16_t* ell_idx, float* ell_val, uint16_t* ell_nnz, int32_t ell_width,
                  int32_t* ell_overflow, flc_stream_t stream);

/* ------------------------------------------------------------------ a5: buckets
 * A.2 bucket rule: round(((mz - 1.00794) * max(|z|,1)) / 1.0005079) // mz_interval,
 * spectra ordered by (charge, interval, precursor m/z) (stable).
 *   order      [n] int32   sorted position -> input index
 *   key_sorted [n] uint32  charge << 24 | interval, in sorted order
 *   mz_sorted  [n] float64 precursor m/z in sorted order
 *   bucket_ptr [n + 1] int64, first *n_buckets + 1 entries valid
 * With a host pointer for n_buckets the stream is synchronised and the count returned; with NULL nothing is
 * read back and the count is left in *n_buckets_dev (see DESIGN.md section 5 for how callers use that). */
FLC_API size_t flc_bucket_sort_workspace_bytes(int64_t n);
FLC_API int flc_bucket_sort(const double* precursor_mz, const int32_t* charge, int64_t n,
                    int32_t mz_interval,
                    int32_t* order, uint32_t* key_sorted, double* mz_sorted,
                    int64_t* bucket_ptr, int64_t* n_buckets /*host; NULL = do not synchronise*/,
                    int64_t pad_buckets /*bucket_ptr[n_buckets .. pad_buckets] = n: empty buckets up to a
                                          host-side upper bound of the count (0 = none; <= n)*/,
                    int64_t* n_buckets_dev /*device, nullable*/,
                    void* workspace, size_t workspace_bytes, flc_stream_t stream);
/* out[i] = in[order[i]] for 4- or 8-byte elements. */
FLC_API int flc_gather(const void* in, const int32_t* order, int64_t n, int elem_bytes,
               void* out, flc_stream_t stream);
/* out[order[i]] = in[i] for 4-byte elements; in == NULL scatters i itself
 * (out = inverse permutation of order). */
FLC_API int flc_scatter32(const void* in, const int32_t* order, int64_t n, void* out,
                  flc_stream_t stream);

/* ------------------------------------------------------------------ 8e: label gather over NVLink peer memory
 * The path's only exchange (running label offset of falcon/falcon.py:189-193).
 * A slot = [n, n_clusters, 0, 0, labels ...] (int32).  flc_scatter_labels_peers writes a slot at element
 * offset slot_offset of each of the `world` buffers in peer_buffers (HOST array of device pointers) with
 * labels in input order: out[order[i]] = labels[i] (order == NULL: out[i]).  The product passes only the
 * rank's own symmetric buffer (world = 1: local stores); n_clusters_dev (device int64, nullable)
 * overrides n_clusters. */
FLC_API int flc_scatter_labels_peers(const int32_t* labels, const int32_t* order, int64_t n,
                             const int64_t* n_clusters_dev, int64_t n_clusters, void* const* peer_buffers,
                             int world, int64_t slot_offset, flc_stream_t stream);
/* slots: HOST array of `world` device pointers, entry r = rank r's slot as mapped into this rank (peer
 * memory: the kernel pulls it with 16-byte loads); slots 16-byte aligned, max_len % 4 == 0 and every slot
 * holds at least 4 + max_len int32.  out[r, i] = slot r's label i plus the cluster counts of slots < r
 * (-1 stays -1, also beyond slot r's length); lens[r] = its length. */
FLC_API int flc_relabel_gathered(const void* const* slots, int world, int64_t max_len, int32_t* out /*[world, max_len]*/,
                         int64_t* lens /*[world]*/, flc_stream_t stream);

/* ------------------------------------------------------------------ a6: IVF train / assign
 * A.2 (faiss IndexIVFFlat over IndexFlatIP): per bucket
 * n_list = 0 (flat) if n < 100 else 2^floor(log2(n/39)) (..., see flc_ivf_plan),
 * spherical k-means (niter iterations, stops early at a fixed point), assignment
 * = arg-max inner product.  Training reads the sparse (ELL) rows of flc_vectorize:
 * buckets whose rows fit in shared memory train in one fused kernel, larger ones
 * in a tiled multi-launch path whose assignment runs on tcgen05 tensor cores when
 * x_bf16 (the unit-norm bf16 rows of flc_vectorize) is given -- from the sparse rows
 * expanded in shared memory for buckets of up to 64 lists, from the bf16 rows
 * otherwise -- with every close call re-scored exactly; list sums are 2^-40 fixed
 * point (int64), so all schedules give the same bits (csrc/kmeans.cu, kmeans_tc.cu).
 * Environment switches for A/B runs and tests: FLC_KMEANS_FORCE_TILED,
 * FLC_KMEANS_NO_TC, FLC_KMEANS_TC_DENSE, FLC_KMEANS_SIMT_SMALL, FLC_KMEANS_TIMING. */
/*  nlist[b], nprobe[b] (int32) and centroid_ptr[b] (int64 exclusive scan of nlist)
 *  for every bucket; exhaustive != 0 lifts the nprobe cap (nprobe = nlist).
 *  Synchronises the stream and returns the totals on the host -- unless the host pointers are NULL (below). */
FLC_API int flc_ivf_plan(const int64_t* bucket_ptr, int64_t n_buckets, int32_t n_probe,
                 int exhaustive, int32_t* nlist, int32_t* nprobe,
                 int64_t* centroid_ptr /*[n_buckets+3]: scan, total, max nprobe, max IVF bucket*/,
                 int64_t* total_centroids /*host*/, int32_t* max_nprobe /*host*/,
                 int64_t* max_ivf_bucket /*host, nullable*/, flc_stream_t stream);
/*  (total_centroids == max_nprobe == NULL: no synchronisation; the three totals stay in
 *  centroid_ptr[n_buckets .. n_buckets + 2].) */
FLC_API size_t flc_kmeans_workspace_bytes(int64_t n, int64_t n_buckets, int64_t total_centroids,
                                  int64_t max_ivf_bucket, int32_t ell_width, uint32_t low_dim);
/*  Whether flc_kmeans_train runs its tiled multi-launch path when the largest IVF bucket has
 *  max_ivf_bucket rows (monotone in max_ivf_bucket).  For callers that pass flc_kmeans_train an upper
 *  bound instead of the value read back from flc_ivf_plan. */
FLC_API int flc_kmeans_needs_tiled(int64_t n, int64_t max_ivf_bucket, int32_t ell_width, uint32_t low_dim);
FLC_API int flc_kmeans_train(const uint16_t* ell_idx, const float* ell_val, const uint16_t* ell_nnz,
                     int32_t ell_width,
                     const uint16_t* x_bf16, int64_t ld_bf16 /*nullable: bf16 copy of the (unit-norm) rows; lets
                       large buckets take their assignment from tcgen05 tensor cores (exact: close calls are
                       re-scored in float32)*/,
                     int64_t n, uint32_t low_dim,
                     const int64_t* bucket_ptr, int64_t n_buckets,
                     const int32_t* nlist, const int64_t* centroid_ptr,
                     int64_t total_centroids, int64_t max_ivf_bucket /*from flc_ivf_plan, 0 = unknown*/,
                     int niter,
                     float* centroids /*[total_centroids, low_dim]*/,
                     /* optional final assignment (same outputs as flc_ivf_assign; both or neither): */
                     const int32_t* nprobe, int32_t max_nprobe, int32_t* list_id, int32_t* probes,
                     void* workspace, size_t workspace_bytes, flc_stream_t stream);
/* Debug aid: per-phase cycle counters of the fused trainer (FLC_KMEANS_TIMING=1). */
FLC_API int flc_debug_kmeans_timing(unsigned long long* out16);
/*  list_id[i] (int32, bucket-local list of row i, 0 for flat buckets) and
 *  probes[i * max_nprobe + j] (int32 list ids best first, -1 padded), both from
 *  float64 inner products with ties to the lower id. */
FLC_API int flc_ivf_assign(const float* x, int64_t ld, int64_t n, uint32_t low_dim,
                   const int64_t* bucket_ptr, int64_t n_buckets,
                   const int32_t* nlist, const int32_t* nprobe,
                   const int64_t* centroid_ptr, const float* centroids,
                   int32_t max_nprobe,
                   const uint16_t* ell_idx, const float* ell_val, int32_t ell_width /*nullable ELL copy*/,
                   int32_t* list_id, int32_t* probes,
                   flc_stream_t stream);

/* ------------------------------------------------------------------ a7: inverted-list scan
 * A.2 index.search: bf16 inner products of every query against the members
 * of its bucket (exhaustive) or of the lists it probes, on tcgen05 tensor
 * cores; (query, candidate) pairs with ip >= threshold are appended to
 * `pairs` (uint64: query << 32 | candidate, global bucket-order row numbers).
 *   impl: 0 = tcgen05/TMA kernel, 1 = SIMT verification kernel (same bf16
 *         inputs, fp32 accumulation on CUDA cores; for bring-up and tests).
 *   list_id/probes NULL: exhaustive within the bucket.
 * The product X X^T is symmetric: the tensor-core kernel only multiplies the
 * tiles at or above the diagonal and appends both (q, c) and (c, q) for every
 * hit with c > q.  Pairs arrive in no particular order.
 * *pair_count (device uint64) receives the number of pairs produced (may
 * exceed pair_capacity: then FLC_ERR_CAPACITY is reported by flc_knn_csr). */
FLC_API size_t flc_scan_workspace_bytes(int64_t n, int64_t n_buckets);
FLC_API int flc_scan_pairs(const uint16_t* x_bf16, int64_t ld_bf16, int64_t n, uint32_t low_dim,
                   const int64_t* bucket_ptr, int64_t n_buckets,
                   const int32_t* list_id, const int32_t* probes, int32_t max_nprobe,
                   const int32_t* nlist,
                   float threshold, int impl,
                   uint64_t* pairs, uint64_t pair_capacity, uint64_t* pair_count,
                   void* workspace, size_t workspace_bytes, flc_stream_t stream);

/* ------------------------------------------------------------------ a7 (tail) + a8 + a9: top-k, filter, CSR
 * For every query: exact re-score of its candidate pairs (float64 accumulate,
 * rounded once to float32), IVF membership check, optional eps cut
 * (dist = max(1 - ip, 0) <= eps; pass NaN to disable), order by (ip desc, id asc),
 * keep n_neighbors_ann, precursor (Da: |d| < tol; ppm: |d| / mz_cand * 1e6 < tol)
 * and RT (|d| < rt_tol, rt == NULL or rt_tol < 0: off) filter in that order,
 * keep n_neighbors, write CSR (float32 dist, int32 column, int64 indptr).
 * Synchronises the stream and returns nnz on the host -- unless nnz is NULL (below). */
FLC_API size_t flc_knn_csr_workspace_bytes(int64_t n, uint64_t n_pairs);
FLC_API int flc_knn_csr(const uint64_t* pairs, const uint64_t* pair_count, uint64_t pair_capacity,
                const float* x, int64_t ld /*dense rows, nullable if ELL given*/,
                const uint16_t* ell_idx, const float* ell_val, int32_t ell_width /*nullable ELL copy*/,
                int64_t n, uint32_t low_dim,
                const double* precursor_mz, const float* rt,
                const int32_t* list_id, const int32_t* probes, int32_t max_nprobe,
                double tol, int tol_mode, double rt_tol,
                int32_t n_neighbors, int32_t n_neighbors_ann, float eps_cut,
                const uint8_t* query_mask /*nullable [n]: rows with a zero byte are left empty (their pairs are
                                            dropped before the re-score) -- callers that need only some rows*/,
                float* dist, int32_t* indices, uint64_t nnz_capacity, int64_t* indptr /*[n+1]*/,
                int64_t* nnz /*host; NULL = do not synchronise: nnz stays in indptr[n], the caller checks
                               *pair_count <= pair_capacity itself, and nnz_capacity must be at least
                               min(pair_capacity, n * n_neighbors)*/,
                void* workspace, size_t workspace_bytes, flc_stream_t stream);

/* ------------------------------------------------------------------ a10: DBSCAN
 * A.4 + sklearn dbscan_inner semantics (sklearn/cluster/_dbscan_inner.pyx:11-41):
 * neighbourhood = row entries with dist <= eps, core = |neighbourhood| >= min_samples,
 * label = rank of the minimum-index core point the point is reachable from
 * along edges leaving core points; -1 = noise.  Returns the number of clusters on the host
 * (synchronises) and / or on the device. */
FLC_API size_t flc_dbscan_workspace_bytes(int64_t n);
FLC_API int flc_dbscan(const float* dist, const int32_t* indices, const int64_t* indptr, int64_t n,
               float eps, int32_t min_samples, int32_t* labels,
               int64_t* n_clusters /*host; NULL = do not synchronise*/, int64_t* n_clusters_dev /*device, nullable*/,
               int32_t n_sweeps /*0: propagate until a sweep changes nothing (the host looks at a flag after
                                  each sweep); > 0: exactly this many sweeps and no look -- then
                                  *unsettled_dev != 0 afterwards means they were not enough*/,
               int32_t* sweeps_used /*host, nullable: sweeps the n_sweeps = 0 loop launched*/,
               int32_t* unsettled_dev /*device int32, required with n_sweeps > 0*/,
               void* workspace, size_t workspace_bytes, flc_stream_t stream);

/* ------------------------------------------------------------------ a11-a15: precursor split
 * falcon/cluster/cluster.py:334-509: inside every DBSCAN cluster (labels_in[i] in [0, n) or -1), 1-D complete
 * linkage on precursor m/z cut at tol (inclusive); sub-clusters with fewer than
 * min_samples members become noise; labels renumbered consecutively.
 * values_sorted != 0 promises precursor_mz ascending inside every cluster in
 * row order (true for bucket-sorted rows) and skips the m/z sort.
 * rt_tol >= 0 adds the retention-time cut of cluster.py:418-429: a second 1-D
 * complete linkage on rt (plain differences, cut at rt_tol), combined with the
 * m/z cut exactly as the reference does -- np.unique(a_mz * 2 + a_rt * 3) on
 * scipy fcluster's ids, a map that is not injective, so the ids are reproduced
 * (depth-first numbering of the dendrogram, scipy _hierarchy.pyx cluster_monocrit).
 * rt: float64 [n] (the values the caller holds, like precursor_mz), required when
 * rt_tol >= 0, else ignored / NULL; pass rt_tol < 0 for "None".
 * Returns #clusters on the host (synchronises) and / or on the device. */
FLC_API size_t flc_split_workspace_bytes(int64_t n, int with_rt);
FLC_API int flc_split_clusters(const int32_t* labels_in, const double* precursor_mz, const double* rt,
                       int64_t n, double tol, int tol_mode, double rt_tol, int32_t min_samples,
                       int values_sorted, int32_t* labels_out,
                       int64_t* n_clusters /*host; NULL = do not synchronise*/,
                       int64_t* n_clusters_dev /*device, nullable*/,
                       void* workspace, size_t workspace_bytes, flc_stream_t stream);

/* ------------------------------------------------------------------ a16: representatives
 * Published falcon get_cluster_representatives (dense descendant at
 * falcon/cluster/cluster.py:512-553): per cluster the member with the smallest
 * mean distance to the members present in its sparse row (rows holding no more
 * than a quarter of the cluster are not eligible; <= 2 members, ties and "no
 * eligible row" give the first member).  labels[i] in [0, n_clusters) or -1 (noise),
 * in the row order of the matrix; medoids[l] = row index (-1 for an unused label). */
FLC_API size_t flc_medoids_workspace_bytes(int64_t n_clusters);
FLC_API int flc_medoids(const float* dist, const int32_t* indices, const int64_t* indptr, int64_t n,
                const int32_t* labels, int64_t n_clusters, int32_t* medoids,
                void* workspace, size_t workspace_bytes, flc_stream_t stream);

/* ------------------------------------------------------------------ preprocessing
 * falcon/cluster/spectrum.py:73-169 (process_spectrum on spectrum_utils 0.3.5): m/z
 * window [mz_min, mz_max] (NaN = open), removal of the precursor peak of every charge
 * state within remove_precursor_tol Da (< 0 = keep), base-peak intensity threshold
 * (min_intensity < 0 = none) and top max_peaks_used cut (0 = none), intensity scaling
 * (0 none, 1 root, 2 log2(1 + x), 3 rank) and L2 normalisation; a spectrum must keep
 * >= min_peaks peaks spanning >= min_mz_range after every step (spectrum.py:27-52).
 * Input peaks ascending in m/z per spectrum.  Outputs: compact CSR of the survivors
 * (out_mz / out_intensity need room for n_peaks, out_indptr [n + 1]; a rejected
 * spectrum keeps zero peaks) and valid[n].  Synchronises; returns the number of
 * surviving peaks on the host. */
FLC_API size_t flc_preprocess_workspace_bytes(int64_t n, int64_t n_peaks);
FLC_API int flc_preprocess(const float* mz, const float* intensity, const int64_t* indptr, int64_t n,
                   int64_t n_peaks, const double* precursor_mz, const int32_t* charge /*nullable: 1*/,
                   int32_t min_peaks, float min_mz_range, float mz_min, float mz_max,
                   float remove_precursor_tol, float min_intensity, int32_t max_peaks_used, int scaling,
                   float* out_mz, float* out_intensity, int64_t* out_indptr, uint8_t* valid,
                   int64_t* n_out_peaks /*host*/, void* workspace, size_t workspace_bytes,
                   flc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FALCON_B200_H_ */
