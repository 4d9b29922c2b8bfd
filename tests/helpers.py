"""Shared helpers for the parity tests: run the oracle and the CUDA path on the
same seeded synthetic spectra."""
from __future__ import annotations

import functools

import numpy as np

from falcon_b200 import synth
from oracle import dbscan as odb
from oracle import ivf as oivf
from oracle import vectorize as ovec

TOL, MODE, EPS = 20.0, "ppm", 0.1


@functools.lru_cache(maxsize=8)
def dataset(n: int, seed: int = 42, lo: float = 700.0, hi: float = 3500.0):
    return synth.generate(n, seed, mass_range=(lo, hi))


def to_device(spectra, dev):
    import torch

    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    return dict(mz=up(spectra.mz), intensity=up(spectra.intensity), indptr=up(spectra.indptr),
                precursor_mz=up(spectra.precursor_mz), charge=up(spectra.precursor_charge),
                rt=up(spectra.retention_time))


def oracle_vectors(spectra, low_dim=400, fragment_tol=0.05, min_mz=101.0, max_mz=1500.0, **kw):
    vec_len, lo, _ = ovec.get_dim(min_mz, max_mz, fragment_tol)
    return ovec.to_vector(spectra.mz, spectra.intensity, spectra.indptr, lo, fragment_tol, vec_len, low_dim, **kw)


def oracle_pipeline(spectra, exhaustive, centroids=None, vectors=None, low_dim=400, mz_interval=1,
                    tol=TOL, mode=MODE, eps=EPS, n_probe=32, rt_tol=None):
    """Oracle run on bucket-sorted spectra.  Returns dict with order, bucket_ptr,
    vectors, csr (full), csr_cut, labels (bucket order)."""
    order, bptr, keys = oivf.bucket_sort(spectra.precursor_mz, spectra.precursor_charge, mz_interval)
    s2 = spectra.take(order)
    x = oracle_vectors(s2, low_dim) if vectors is None else vectors
    rts = None if rt_tol is None else np.asarray(s2.retention_time, np.float32)
    mat, cents = oivf.compute_pairwise_distances(
        x, s2.precursor_mz, rts, bptr, tol, mode, rt_tol, 64, 128, n_probe, exhaustive, centroids)
    cut = oivf.eps_cut(mat, eps)
    labels = odb.generate_clusters(cut.data, cut.indices, cut.indptr, eps, s2.precursor_mz,
                                   None if rts is None else rts.astype(np.float64), tol, mode, rt_tol)
    return dict(order=order, bucket_ptr=bptr, keys=keys, sorted=s2, x=x, csr=mat, csr_cut=cut,
                labels=labels, centroids=cents)


def rows_equal(indptr_a, idx_a, indptr_b, idx_b):
    """Row-wise equality of two CSR index structures (order inside rows matters)."""
    return np.array_equal(indptr_a, indptr_b) and np.array_equal(idx_a, idx_b)
