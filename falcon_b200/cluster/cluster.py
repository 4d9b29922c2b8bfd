"""Pairwise distances and cluster generation -- mirror of ``falcon.cluster.cluster``.

``compute_pairwise_distances`` and ``generate_clusters`` keep the names,
argument order and meaning of published falcon 0.1.x (the
surviving pieces are at /root/reference/falcon/cluster/cluster.py:24-156,
334-509).  Results are numpy / scipy objects like the reference's; the work is
done by the CUDA kernels behind ``include/falcon_b200.h``.
"""
from __future__ import annotations

import functools
import logging
import pickle
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.sparse as ss
import torch

from .. import pipeline, synth

logger = logging.getLogger("falcon")


def _load_bucket(item, process_spectrum: Optional[Callable]) -> List[dict]:
    if isinstance(item, (str, bytes)):
        try:
            import joblib

            spectra = joblib.load(item)
        except Exception:  # plain pickle
            with open(item, "rb") as fh:
                spectra = pickle.load(fh)
    else:
        spectra = list(item)
    if process_spectrum is not None:
        spectra = [s for s in (process_spectrum(s) for s in spectra) if s is not None]
    return spectra


def _settings_from(vectorize, **kw) -> pipeline.Settings:
    s = pipeline.Settings(**kw)
    if isinstance(vectorize, functools.partial):
        k = vectorize.keywords
        if "dim" in k:
            s.low_dim = int(k["dim"])
        if "bin_size" in k:
            s.fragment_tol = float(k["bin_size"])
    return s


def compute_pairwise_distances(
    n_spectra: int,
    bucket_filenames: Union[Sequence, synth.SpectrumSet],
    process_spectrum: Optional[Callable],
    vectorize: Optional[Callable],
    precursor_tol_mass: float,
    precursor_tol_mode: str,
    rt_tol: Optional[float],
    n_neighbors: int,
    n_neighbors_ann: int,
    batch_size: int,
    n_probe: int,
    *,
    eps: Optional[float] = 0.1,
    exhaustive: bool = False,
    mz_interval: int = 1,
    min_mz: float = 101.0,
    max_mz: float = 1500.0,
) -> Tuple[ss.csr_matrix, "object"]:
    """Sparse k-NN cosine-distance matrix of all spectra of one charge.

    ``bucket_filenames``: the per-bucket spectrum files of the reference
    (joblib/pickle lists of spectrum dicts), in-memory lists of spectrum dicts,
    or one ``SpectrumSet`` (buckets are then derived with the reference's bucket
    rule).  Rows of the result follow the bucket order, sorted by precursor m/z
    inside a bucket, exactly as the reference concatenates its buckets; the
    returned metadata (``pandas.DataFrame``) lists the spectra in that order.

    ``eps`` (keyword): when given, only entries with ``dist <= eps`` are
    produced -- the part of the matrix ``generate_clusters`` reads (the eps cut
    is fused into the scan); ``eps=None`` returns every
    neighbour like the reference.
    """
    import pandas as pd

    if precursor_tol_mode not in ("Da", "ppm"):
        raise ValueError("Unknown precursor tolerance mode")
    if n_neighbors_ann < n_neighbors:
        raise ValueError("n_neighbors_ann should be equal or greater than n_neighbors")
    if isinstance(bucket_filenames, synth.SpectrumSet):
        spectra = bucket_filenames
        idents = [f"spectrum:{i}" for i in range(len(spectra))]
    else:
        dicts: List[dict] = []
        for item in bucket_filenames:
            dicts.extend(_load_bucket(item, process_spectrum))
        spectra = synth.SpectrumSet.from_dicts(dicts)
        idents = [d.get("identifier", str(i)) for i, d in enumerate(dicts)]
    if len(spectra) != n_spectra:
        raise ValueError(f"n_spectra = {n_spectra} but the buckets hold {len(spectra)} spectra")
    settings = _settings_from(
        vectorize, precursor_tol_mass=precursor_tol_mass, precursor_tol_mode=precursor_tol_mode,
        rt_tol=rt_tol, n_neighbors=n_neighbors, n_neighbors_ann=n_neighbors_ann, batch_size=batch_size,
        n_probe=n_probe, eps=0.0 if eps is None else float(eps), eps_cut=eps is not None,
        exhaustive=exhaustive, mz_interval=mz_interval, min_mz=min_mz, max_mz=max_mz)
    hp = pipeline.HotPath(settings)
    dev = hp.device
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    n = len(spectra)
    if n == 0:
        return ss.csr_matrix((0, 0), dtype=np.float32), pd.DataFrame(
            {"identifier": [], "precursor_charge": [], "precursor_mz": [], "retention_time": []})
    pmz = up(spectra.precursor_mz, np.float64)
    z = up(spectra.precursor_charge, np.int32)
    rt = up(spectra.retention_time, np.float32)
    buckets = hp.bucket_sort(pmz, z, rt)
    v = hp.vectorize(up(spectra.mz, np.float32), up(spectra.intensity, np.float32),
                     up(spectra.indptr, np.int64), buckets.order)
    ivf = None if exhaustive else hp.build_ivf(v, buckets)
    order = buckets.order.cpu().numpy()
    idx_dtype = np.int32 if n * n_neighbors < 2 ** 31 else np.int64
    mat = ss.csr_matrix((n, n), dtype=np.float32)
    if eps is not None:
        g = hp.knn_graph(v, buckets, ivf)
        mat.data = g.dist.cpu().numpy()
        mat.indices = g.indices.cpu().numpy().astype(idx_dtype, copy=False)
        mat.indptr = g.indptr.cpu().numpy().astype(idx_dtype, copy=False)
    else:
        # the full matrix: every bucket yields rows^2 candidate pairs, so it is built range by range
        data, indices, counts = [], [], []
        for r0, r1, g in hp.uncut_graph_ranges(v, buckets, ivf):
            ptr_ = g.indptr[r0: r1 + 1].cpu().numpy()
            data.append(g.dist[int(ptr_[0]): int(ptr_[-1])].cpu().numpy())
            indices.append(g.indices[int(ptr_[0]): int(ptr_[-1])].cpu().numpy())
            counts.append(np.diff(ptr_))
        indptr = np.zeros(n + 1, np.int64)
        np.cumsum(np.concatenate(counts), out=indptr[1:])
        mat.data = np.concatenate(data)
        mat.indices = np.concatenate(indices).astype(idx_dtype, copy=False)
        mat.indptr = indptr.astype(idx_dtype, copy=False)
    metadata = pd.DataFrame({
        "identifier": [idents[i] for i in order],
        "precursor_charge": spectra.precursor_charge[order],
        "precursor_mz": spectra.precursor_mz[order],
        "retention_time": spectra.retention_time[order],
    })
    return mat, metadata


def generate_clusters(
    pairwise_dist_matrix: ss.csr_matrix,
    eps: float,
    precursor_mzs: np.ndarray,
    rts: Optional[np.ndarray],
    precursor_tol_mass: float,
    precursor_tol_mode: str,
    rt_tol: Optional[float] = None,
) -> np.ndarray:
    """DBSCAN clustering of the pairwise distance matrix; noise = -1.

    ``min_samples = 2`` (/root/reference/falcon/cluster/cluster.py:66); clusters
    are then split so that no cluster exceeds the precursor tolerance
    (cluster.py:334-509) -- nor, with ``rt_tol``, the retention-time tolerance
    (cluster.py:418-429) -- and relabelled consecutively.
    """
    if precursor_tol_mode not in ("Da", "ppm"):
        raise ValueError("Unknown precursor tolerance mode")
    n = pairwise_dist_matrix.shape[0]
    if len(precursor_mzs) != n:
        raise ValueError("precursor_mzs does not match the distance matrix")
    if rt_tol is not None and (rts is None or len(rts) != n):
        raise ValueError("rt_tol is set but rts does not match the distance matrix")
    if n == 0:
        return np.zeros(0, np.int64)
    settings = pipeline.Settings(eps=float(eps), precursor_tol_mass=precursor_tol_mass,
                                 precursor_tol_mode=precursor_tol_mode, rt_tol=rt_tol)
    hp = pipeline.HotPath(settings)
    dev = hp.device
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    g = pipeline.KnnGraph(up(pairwise_dist_matrix.data, np.float32), up(pairwise_dist_matrix.indices, np.int32),
                          up(pairwise_dist_matrix.indptr, np.int64), int(pairwise_dist_matrix.nnz), 0)
    labels, _ = hp.dbscan(g, n)
    out, n_clusters = hp.split(labels, up(precursor_mzs, np.float64), values_sorted=False,
                               rt=up(rts, np.float64) if rt_tol is not None else None)
    out = out.cpu().numpy().astype(np.int64)
    logger.info("%d spectra grouped in %d clusters, %d spectra remain as singletons",
                int((out != -1).sum()), n_clusters, int((out == -1).sum()))
    return out


def get_cluster_representatives(
    clusters: np.ndarray,
    pairwise_indptr: np.ndarray,
    pairwise_indices: np.ndarray,
    pairwise_data: np.ndarray,
) -> Optional[np.ndarray]:
    """Indexes of the cluster representative spectra (medoids) -- published falcon
    ``get_cluster_representatives`` (the snapshot's dense descendant is
    /root/reference/falcon/cluster/cluster.py:512-553).

    ``clusters``: label of every row of the pairwise distance matrix (-1 = noise).
    Returns, for every distinct non-noise label in ascending order, the row whose
    mean distance to the cluster members present in its sparse row is smallest;
    ``None`` if there is no cluster.
    """
    clusters = np.asarray(clusters)
    n = clusters.shape[0]
    if len(pairwise_indptr) != n + 1:
        raise ValueError("clusters does not match the distance matrix")
    uniq, dense = np.unique(clusters[clusters >= 0], return_inverse=True)
    if uniq.size == 0:
        return None
    lab = np.full(n, -1, np.int32)
    lab[clusters >= 0] = dense.astype(np.int32)
    hp = pipeline.HotPath(pipeline.Settings())
    dev = hp.device
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    g = pipeline.KnnGraph(up(pairwise_data, np.float32), up(pairwise_indices, np.int32),
                          up(pairwise_indptr, np.int64), int(len(pairwise_data)), 0)
    return hp.medoids(g, up(lab, np.int32), int(uniq.size)).cpu().numpy()
