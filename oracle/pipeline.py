"""End-to-end CPU run of the oracle -- the checker for whole-path parity and the
timed CPU baseline of ``bench.py`` (``cpu_baseline`` / ``--impl reference``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  "CPU restatement
(faiss-cpu unavailable offline)": numpy/OpenBLAS ``sgemm`` per bucket stands in
for faiss' scanner, sklearn's own ``dbscan_inner`` is the DBSCAN.  Buckets are
independent, so the search stage is spread over ``n_jobs`` worker processes
(joblib), the way the reference spreads its work with joblib
(/root/reference/falcon/cluster/cluster.py:115-117); DBSCAN is serial as in
sklearn.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as ss

from . import dbscan as odb
from . import ivf as oivf
from . import vectorize as ovec


def _search_range(args):
    (mz, inten, indptr, pmz, bptr, vec_len, min_mz, p) = args
    x = ovec.to_vector(mz, inten, indptr, min_mz, p["fragment_tol"], vec_len, p["low_dim"])
    mat, _ = oivf.compute_pairwise_distances(
        x, pmz, None, bptr, p["tol"], p["mode"], None, p["n_neighbors"], p["n_neighbors_ann"],
        p["n_probe"], p["exhaustive"], None, use_f32_gemm=p["f32_gemm"])
    return mat.data, mat.indices, np.diff(mat.indptr)


def run(spectra, *, low_dim=400, fragment_tol=0.05, min_mz=101.0, max_mz=1500.0, tol=20.0, mode="ppm",
        eps=0.1, n_neighbors=64, n_neighbors_ann=128, n_probe=32, exhaustive=False, mz_interval=1,
        n_jobs=1, f32_gemm=True):
    """Cluster ``spectra`` (a SpectrumSet-like object).  Returns labels in input
    order and per-stage wall times."""
    t = {}
    t0 = time.perf_counter()
    vec_len, lo, _ = ovec.get_dim(min_mz, max_mz, fragment_tol)
    order, bptr, _ = oivf.bucket_sort(spectra.precursor_mz, spectra.precursor_charge, mz_interval)
    s2 = spectra.take(order)
    t["bucket_sort"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    p = dict(low_dim=low_dim, fragment_tol=fragment_tol, tol=tol, mode=mode, n_neighbors=n_neighbors,
             n_neighbors_ann=n_neighbors_ann, n_probe=n_probe, exhaustive=exhaustive, f32_gemm=f32_gemm)
    n = len(s2)
    nb = bptr.shape[0] - 1
    n_chunks = max(1, min(nb, n_jobs * 4))
    cuts = np.unique(np.linspace(0, nb, n_chunks + 1).astype(np.int64))
    jobs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        r0, r1 = int(bptr[a]), int(bptr[b])
        p0, p1 = int(s2.indptr[r0]), int(s2.indptr[r1])
        jobs.append((s2.mz[p0:p1], s2.intensity[p0:p1], s2.indptr[r0: r1 + 1] - p0, s2.precursor_mz[r0:r1],
                     bptr[a: b + 1] - r0, vec_len, lo, p))
    if n_jobs > 1 and len(jobs) > 1:
        import joblib

        parts = joblib.Parallel(n_jobs=n_jobs)(joblib.delayed(_search_range)(j) for j in jobs)
    else:
        parts = [_search_range(j) for j in jobs]
    data = np.concatenate([q[0] for q in parts]) if parts else np.zeros(0, np.float32)
    indices = np.concatenate([q[1] + int(bptr[a]) for q, a in zip(parts, cuts[:-1])]) if parts else np.zeros(0, np.int64)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.concatenate([q[2] for q in parts]) if parts else np.zeros(0, np.int64), out=indptr[1:])
    t["vectorize+pairwise"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    labels_sorted = odb.generate_clusters(data, indices, indptr, eps, s2.precursor_mz, None, tol, mode)
    t["generate_clusters"] = time.perf_counter() - t0
    labels = np.empty(n, np.int64)
    labels[order] = labels_sorted
    mat = ss.csr_matrix((n, n), dtype=np.float32)
    mat.data, mat.indices, mat.indptr = data, indices, indptr
    return labels, t, mat
