"""Bucketed inverted-file nearest-neighbour search oracle.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  **Parity unpinned**: the
code this restates (``compute_pairwise_distances`` of published falcon 0.1.x on
top of faiss ``IndexIVFFlat``/``IndexFlatIP``) is neither in the mounted
snapshot nor installable here (faiss absent); it follows SURVEY.md Appendix
A.2 and the README prose (/root/reference/README.md:132-142).  The exhaustive
mode is self-checking against a brute-force ``X @ X.T``.

Conventions this oracle fixes where faiss leaves them open (the CUDA path
follows the same, so results are comparable bit for bit):
  * inner products are accumulated in float64 and rounded once to float32
    (order independent; faiss' float32 SIMD sum differs by ~1e-7);
  * ties in inner product are broken by the lower row id (faiss: unspecified);
  * coarse assignment and probe selection use float64 inner products with
    ties to the lower centroid id;
  * k-means initialisation is evenly strided rows instead of faiss' seeded
    random permutation (north_star compares with *shared* centroids).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as ss

H_MASS = 1.00794
ISOTOPE = 1.0005079


# --------------------------------------------------------------------------- buckets
def bucket_ids(precursor_mz, charge, mz_interval: int = 1) -> np.ndarray:
    """Precursor-mass interval of every spectrum:
    ``round(((mz - 1.00794) * max(|z|, 1)) / 1.0005079) // mz_interval``
    (Python ``round`` = round-half-even)."""
    z = np.maximum(np.abs(np.asarray(charge, np.int64)), 1)
    x = (np.asarray(precursor_mz, np.float64) - H_MASS) * z.astype(np.float64) / ISOTOPE
    return (np.rint(x).astype(np.int64) // int(mz_interval)).astype(np.int64)


def bucket_sort(precursor_mz, charge, mz_interval: int = 1):
    """Order spectra by (charge, interval, precursor m/z); stable.

    Returns ``order`` (sorted position -> input index), ``bucket_ptr``
    (``int64[n_buckets + 1]`` offsets into the sorted order) and the
    ``uint32`` bucket key of every bucket (``charge << 24 | interval``).
    """
    interval = bucket_ids(precursor_mz, charge, mz_interval)
    z = np.clip(np.asarray(charge, np.int64), 0, 255)
    key = (z << 24) | (interval & 0xFFFFFF)
    order = np.lexsort((np.asarray(precursor_mz, np.float64), key)).astype(np.int64)
    ks = key[order]
    heads = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if ks.size else np.zeros(0, np.int64)
    bucket_ptr = np.r_[heads, ks.size].astype(np.int64)
    return order, bucket_ptr, ks[heads].astype(np.uint32) if ks.size else np.zeros(0, np.uint32)


# --------------------------------------------------------------------------- IVF sizing
def n_list_rule(n: int) -> int:
    """faiss index choice of published falcon.  0 = flat index."""
    if n < 100:
        return 0
    if n < 10**6:
        p = 1
        while 39 * p * 2 <= n:  # largest power of two with 39 * p <= n
            p *= 2
        return p
    if n < 10**7:
        return 2**16
    if n < 10**8:
        return 2**18
    return 2**20


def n_probe_rule(n_list: int, n_probe: int, exhaustive: bool = False) -> int:
    """``index.nprobe = min(ceil(n_list / 8), n_probe)``;
    ``exhaustive`` lifts the cap (north_star: n_probe = nlist)."""
    if n_list == 0:
        return 0
    if exhaustive:
        return n_list
    return max(1, min(math.ceil(n_list / 8), int(n_probe)))


# --------------------------------------------------------------------------- k-means
def kmeans_train(x: np.ndarray, n_list: int, n_iter: int = 10) -> np.ndarray:
    """Spherical k-means in the spirit of faiss ``Clustering`` for the IP metric
    (niter = 10, centroids re-normalised every iteration, empty
    clusters split off the largest one with a +-1/1024 perturbation).

    Conventions shared with the CUDA trainers (csrc/kmeans.cu): initialisation
    from rows ``(c * n) // n_list``; float32 argmax assignment (ties -> lower
    id); list sums in 2^-40 fixed point (``sum rint(x * 2^40)`` as int64, exact
    and order independent); mean = float32(sum * (2^-40 / count)); unit norm by
    float32(c * (1 / sqrt(sum c^2))); iterations stop at a fixed point.
    """
    n, d = x.shape
    x = np.ascontiguousarray(x, np.float32)
    cent = x[(np.arange(n_list, dtype=np.int64) * n) // n_list].copy()
    xq = np.rint(x.astype(np.float64) * 2.0 ** 40).astype(np.int64)
    eps = np.float32(1.0 / 1024.0)
    prev = None
    for _ in range(n_iter):
        ip = x @ cent.T
        assign = np.argmax(ip, axis=1)
        counts = np.bincount(assign, minlength=n_list).astype(np.int64)
        nz = counts > 0
        if prev is not None and nz.all() and np.array_equal(assign, prev):
            break
        prev = assign
        sums = np.zeros((n_list, d), np.int64)
        np.add.at(sums, assign, xq)
        new = cent.copy()
        scale = (2.0 ** -40) / counts[nz].astype(np.float64)
        new[nz] = (sums[nz].astype(np.float64) * scale[:, None]).astype(np.float32)
        # Split the largest cluster into every empty one.
        cnt = counts.astype(np.float64)
        sign = np.where(np.arange(d) % 2 == 0, np.float32(1) + eps, np.float32(1) - eps)
        for ci in np.flatnonzero(~nz):
            cj = int(np.argmax(cnt))
            new[ci] = new[cj] * sign
            new[cj] = new[cj] * (np.float32(2) - sign)
            cnt[ci] = cnt[cj] / 2
            cnt[cj] -= cnt[ci]
        ssq = (new.astype(np.float64) ** 2).sum(axis=1, keepdims=True)
        inv = np.where(ssq > 0, 1.0 / np.sqrt(np.where(ssq > 0, ssq, 1.0)), 1.0)
        cent = (new.astype(np.float64) * inv).astype(np.float32)
    return cent


def coarse_ip(x: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    return x.astype(np.float64) @ centroids.astype(np.float64).T


def assign_lists(x: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """List id of every vector: argmax float64 inner product, ties -> lower id."""
    return np.argmax(coarse_ip(x, centroids), axis=1).astype(np.int32)


def probe_lists(x: np.ndarray, centroids: np.ndarray, n_probe: int) -> np.ndarray:
    """``int32[n, n_probe]`` probed list ids per query, best first."""
    ip = coarse_ip(x, centroids)
    order = np.argsort(-ip, axis=1, kind="stable")
    return order[:, :n_probe].astype(np.int32)


# --------------------------------------------------------------------------- search + filter
def exact_ip(xq: np.ndarray, xc: np.ndarray) -> np.ndarray:
    """float64-accumulated inner products rounded once to float32."""
    return (xq.astype(np.float64) @ xc.astype(np.float64).T).astype(np.float32)


def tolerance_mask(mz_q, mz_c, rt_q, rt_c, tol, mode, rt_tol):
    """Precursor / RT compatibility of (query, candidate) pairs:
    Da: ``|dmz| < tol``; ppm: ``|dmz| / mz_candidate * 1e6 < tol`` (strict);
    optional ``|drt| < rt_tol``.  float64 throughout."""
    dm = np.abs(mz_q[:, None] - mz_c[None, :])
    if mode == "Da":
        m = dm < tol
    elif mode == "ppm":
        m = dm / mz_c[None, :] * 10**6 < tol
    else:
        raise ValueError("Unknown precursor tolerance mode")
    if rt_tol is not None:
        m &= np.abs(rt_q.astype(np.float64)[:, None] - rt_c.astype(np.float64)[None, :]) < rt_tol
    return m


def search_bucket(
    x: np.ndarray,
    precursor_mz: np.ndarray,
    rt: np.ndarray | None,
    tol: float,
    mode: str,
    rt_tol: float | None,
    n_neighbors: int,
    n_neighbors_ann: int,
    n_probe: int,
    exhaustive: bool = False,
    centroids: np.ndarray | None = None,
    use_f32_gemm: bool = False,
):
    """One bucket: ANN search, precursor filter, distances.

    Returns ``(dist float32[n, k], idx int32[n, k] (bucket local, -1 pad),
    count int32[n], centroids or None)``.
    """
    n = x.shape[0]
    k_ann = min(n_neighbors_ann, n)
    sim = (x @ x.T).astype(np.float32) if use_f32_gemm else exact_ip(x, x)
    n_list = n_list_rule(n)
    cand = None
    if n_list and not exhaustive:
        if centroids is None:
            centroids = kmeans_train(x, n_list)
        lists = assign_lists(x, centroids)
        probes = probe_lists(x, centroids, n_probe_rule(n_list, n_probe))
        probed = np.zeros((n, n_list), bool)
        np.put_along_axis(probed, probes.astype(np.int64), True, axis=1)
        cand = probed[:, lists]  # [query, candidate]
        sim = np.where(cand, sim, -np.inf).astype(np.float32)
    # top k_ann by inner product, ties -> lower id (stable sort on -sim).
    order = np.argsort(-sim, axis=1, kind="stable")[:, :k_ann]
    top_sim = np.take_along_axis(sim, order, axis=1)
    valid = np.isfinite(top_sim)
    rt_ = None if rt is None else np.asarray(rt)
    ok = tolerance_mask(
        precursor_mz, precursor_mz, rt_, rt_, tol, mode, rt_tol if rt is not None else None
    )
    keep = np.take_along_axis(ok, order, axis=1) & valid
    rank = np.cumsum(keep, axis=1) - 1
    keep &= rank < n_neighbors
    count = keep.sum(axis=1).astype(np.int32)
    k = n_neighbors
    dist = np.zeros((n, k), np.float32)
    idx = np.full((n, k), -1, np.int32)
    r, c = np.nonzero(keep)
    slot = rank[r, c]
    dist[r, slot] = np.maximum(np.float32(1) - top_sim[r, c], np.float32(0))
    idx[r, slot] = order[r, c]
    return dist, idx, count, centroids


def compute_pairwise_distances(
    vectors: np.ndarray,
    precursor_mz: np.ndarray,
    rt: np.ndarray | None,
    bucket_ptr: np.ndarray,
    tol: float,
    mode: str,
    rt_tol: float | None = None,
    n_neighbors: int = 64,
    n_neighbors_ann: int = 128,
    n_probe: int = 32,
    exhaustive: bool = False,
    centroids: list | None = None,
    use_f32_gemm: bool = False,
):
    """Sparse k-NN cosine-distance matrix over bucket-sorted spectra.

    ``vectors`` etc. are already in bucket order (``bucket_sort``).  Returns a
    ``scipy.sparse.csr_matrix`` (float32, N x N, rows in similarity order,
    column ids global) and the list of per-bucket centroids used (None for
    flat buckets).
    """
    n = vectors.shape[0]
    data, indices, counts, cents = [], [], np.zeros(n, np.int64), []
    for b in range(bucket_ptr.shape[0] - 1):
        s, e = int(bucket_ptr[b]), int(bucket_ptr[b + 1])
        c_in = None if centroids is None else centroids[b]
        dist, idx, cnt, c_out = search_bucket(
            vectors[s:e], precursor_mz[s:e], None if rt is None else rt[s:e],
            tol, mode, rt_tol, n_neighbors, n_neighbors_ann, n_probe,
            exhaustive, c_in, use_f32_gemm,
        )
        m = idx >= 0
        data.append(dist[m])
        indices.append(idx[m].astype(np.int64) + s)
        counts[s:e] = cnt
        cents.append(c_out)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    data = np.concatenate(data) if data else np.zeros(0, np.float32)
    indices = np.concatenate(indices) if indices else np.zeros(0, np.int64)
    idx_dtype = np.int32 if n * n_neighbors < 2**31 else np.int64
    mat = ss.csr_matrix((n, n), dtype=np.float32)
    mat.data, mat.indices, mat.indptr = (
        data.astype(np.float32), indices.astype(idx_dtype), indptr.astype(idx_dtype if idx_dtype == np.int64 else np.int64),
    )
    return mat, cents


def eps_cut(mat: ss.csr_matrix, eps: float) -> ss.csr_matrix:
    """Drop entries with ``data > float32(eps)`` keeping row order -- the part of
    the matrix ``generate_clusters`` reads (``mask = data <= eps``)."""
    keep = mat.data <= np.float32(eps)
    rows = np.repeat(np.arange(mat.shape[0]), np.diff(mat.indptr))
    counts = np.bincount(rows[keep], minlength=mat.shape[0])
    indptr = np.zeros(mat.shape[0] + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    out = ss.csr_matrix(mat.shape, dtype=np.float32)
    out.data, out.indices, out.indptr = mat.data[keep], mat.indices[keep], indptr
    return out
