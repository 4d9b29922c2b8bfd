"""Seeded synthetic MS/MS spectra.

The generator produces what ``falcon``'s preprocessing hands to the clustering
hot path: per spectrum a peak list sorted by m/z with unit-L2 float32
intensities and at most ``max_peaks`` peaks (mirrors the output contract of
``process_spectrum``, /root/reference/falcon/cluster/spectrum.py:157-169, and the
``max_peaks_used`` default of 50, /root/reference/falcon/config.py:169-171),
plus precursor m/z, precursor charge and retention time.

Peaks are kept in CSR form (``mz``/``intensity`` float32 of length
``indptr[-1]``, ``indptr`` int64 of length ``n + 1``) because that is the layout
the vectorisation kernel consumes (``_to_vector`` builds the same three arrays,
/root/reference/falcon/cluster/spectrum.py:280-296).

Everything is vectorised numpy so that 1 M spectra take a few seconds.
"""
from __future__ import annotations

import dataclasses

import numpy as np

PROTON = 1.00728


@dataclasses.dataclass
class SpectrumSet:
    """CSR peak arrays + per-spectrum metadata for ``n`` spectra."""

    mz: np.ndarray  # float32 [n_peaks]
    intensity: np.ndarray  # float32 [n_peaks]
    indptr: np.ndarray  # int64 [n + 1]
    precursor_mz: np.ndarray  # float64 [n]
    precursor_charge: np.ndarray  # int32 [n]
    retention_time: np.ndarray  # float32 [n]
    template: np.ndarray | None = None  # int64 [n] ground-truth template id

    def __len__(self) -> int:
        return int(self.indptr.shape[0] - 1)

    @property
    def n_peaks(self) -> int:
        return int(self.indptr[-1])

    def take(self, idx: np.ndarray) -> "SpectrumSet":
        """Sub-set (and re-order) the spectra by index."""
        idx = np.asarray(idx, dtype=np.int64)
        counts = (self.indptr[1:] - self.indptr[:-1])[idx]
        indptr = np.zeros(idx.shape[0] + 1, np.int64)
        np.cumsum(counts, out=indptr[1:])
        # Gather peak ranges.
        starts = self.indptr[:-1][idx]
        pos = np.arange(int(indptr[-1]), dtype=np.int64)
        owner = np.repeat(np.arange(idx.shape[0], dtype=np.int64), counts)
        src = starts[owner] + (pos - indptr[:-1][owner])
        return SpectrumSet(
            self.mz[src],
            self.intensity[src],
            indptr,
            self.precursor_mz[idx],
            self.precursor_charge[idx],
            self.retention_time[idx],
            None if self.template is None else self.template[idx],
        )

    def as_dicts(self) -> list[dict]:
        """Spectra as the dicts ``process_spectrum`` returns
        (/root/reference/falcon/cluster/spectrum.py:161-169)."""
        out = []
        for i in range(len(self)):
            a, b = int(self.indptr[i]), int(self.indptr[i + 1])
            out.append(
                {
                    "identifier": f"synth:{i}",
                    "precursor_mz": float(self.precursor_mz[i]),
                    "precursor_charge": int(self.precursor_charge[i]),
                    "mz": self.mz[a:b],
                    "intensity": self.intensity[a:b],
                    "retention_time": float(self.retention_time[i]),
                    "filename": "synthetic",
                }
            )
        return out

    @staticmethod
    def from_dicts(spectra: list[dict]) -> "SpectrumSet":
        n = len(spectra)
        counts = np.fromiter((len(s["mz"]) for s in spectra), np.int64, n)
        indptr = np.zeros(n + 1, np.int64)
        np.cumsum(counts, out=indptr[1:])
        if n:
            mz = np.concatenate([np.asarray(s["mz"], np.float32) for s in spectra])
            inten = np.concatenate(
                [np.asarray(s["intensity"], np.float32) for s in spectra]
            )
        else:
            mz = np.zeros(0, np.float32)
            inten = np.zeros(0, np.float32)
        rt = np.asarray(
            [
                s.get("retention_time") if s.get("retention_time") is not None else 0.0
                for s in spectra
            ],
            np.float32,
        )
        return SpectrumSet(
            mz,
            inten,
            indptr,
            np.asarray([s["precursor_mz"] for s in spectra], np.float64),
            np.asarray(
                [
                    s["precursor_charge"] if s["precursor_charge"] is not None else 0
                    for s in spectra
                ],
                np.int32,
            ),
            rt,
        )


def concat(sets) -> SpectrumSet:
    """Concatenation of SpectrumSets (templates are dropped: their ids are per set)."""
    sets = list(sets)
    if len(sets) == 1:
        return sets[0]
    cat = np.concatenate
    indptr = np.zeros(sum(len(s) for s in sets) + 1, np.int64)
    np.cumsum(cat([np.diff(s.indptr) for s in sets]), out=indptr[1:])
    return SpectrumSet(cat([s.mz for s in sets]), cat([s.intensity for s in sets]), indptr,
                       cat([s.precursor_mz for s in sets]), cat([s.precursor_charge for s in sets]),
                       cat([s.retention_time for s in sets]))


def generate(
    n: int,
    seed: int = 42,
    *,
    max_peaks: int = 50,
    min_frag_mz: float = 101.0,
    max_frag_mz: float = 1500.0,
    mass_range: tuple[float, float] = (700.0, 3500.0),
    mean_cluster_size: float = 4.0,
    shuffle: bool = True,
) -> SpectrumSet:
    """Generate ``n`` synthetic spectra: peptide templates + noisy replicates.

    ``numpy.random.Generator(PCG64(seed))``; charge 2/3 with p = 0.6/0.4; neutral
    mass ~ U(mass_range); 20..50 fragments per template, m/z ~ U(101, 1500),
    base intensity ~ LogNormal(0, 1); cluster sizes ~ Geometric(1/mean);
    replicate jitter: fragment m/z N(0, 0.005 Da), intensity x LogNormal(0, 0.2),
    peak drop-out 0.1, 0..5 noise peaks at <= 5 % of the base peak, precursor
    jitter U(+-5 ppm). Peaks are sorted by m/z, truncated to the ``max_peaks``
    most intense, and L2-normalised in float32.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if n == 0:
        z = np.zeros(0)
        return SpectrumSet(
            z.astype(np.float32), z.astype(np.float32), np.zeros(1, np.int64),
            z.astype(np.float64), z.astype(np.int32), z.astype(np.float32),
            z.astype(np.int64),
        )
    # Cluster sizes until they cover n spectra.
    n_templ_guess = int(n / mean_cluster_size * 1.3) + 16
    sizes = rng.geometric(1.0 / mean_cluster_size, n_templ_guess).astype(np.int64)
    csum = np.cumsum(sizes)
    while csum[-1] < n:
        more = rng.geometric(1.0 / mean_cluster_size, n_templ_guess).astype(np.int64)
        sizes = np.concatenate([sizes, more])
        csum = np.cumsum(sizes)
    n_templ = int(np.searchsorted(csum, n, "left")) + 1
    sizes = sizes[:n_templ].copy()
    sizes[-1] -= csum[n_templ - 1] - n
    template = np.repeat(np.arange(n_templ, dtype=np.int64), sizes)
    assert template.shape[0] == n

    # Templates.
    t_charge = np.where(rng.random(n_templ) < 0.6, 2, 3).astype(np.int32)
    t_mass = rng.uniform(mass_range[0], mass_range[1], n_templ)
    t_pmz = (t_mass + t_charge * PROTON) / t_charge
    t_rt = rng.uniform(0.0, 7200.0, n_templ)
    t_nfrag = rng.integers(20, max_peaks + 1, n_templ)
    width = max_peaks + 6  # template fragments + up to 5 noise peaks (+1 pad)
    t_fmz = rng.uniform(min_frag_mz, max_frag_mz, (n_templ, max_peaks))
    t_fint = rng.lognormal(0.0, 1.0, (n_templ, max_peaks))
    t_valid = np.arange(max_peaks)[None, :] < t_nfrag[:, None]

    # Replicates, processed in chunks to bound memory.
    mz_out, int_out, cnt_out = [], [], []
    chunk = 1 << 18
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        t = template[s:e]
        m = e - s
        fmz = np.full((m, width), np.inf)
        fint = np.zeros((m, width))
        keep = t_valid[t] & (rng.random((m, max_peaks)) >= 0.1)
        fmz[:, :max_peaks] = np.where(
            keep, t_fmz[t] + rng.normal(0.0, 0.005, (m, max_peaks)), np.inf
        )
        fint[:, :max_peaks] = np.where(
            keep, t_fint[t] * rng.lognormal(0.0, 0.2, (m, max_peaks)), 0.0
        )
        base = fint.max(axis=1, keepdims=True)
        base[base == 0] = 1.0
        n_noise = rng.integers(0, 6, m)
        nz = np.arange(5)[None, :] < n_noise[:, None]
        fmz[:, max_peaks : max_peaks + 5] = np.where(
            nz, rng.uniform(min_frag_mz, max_frag_mz, (m, 5)), np.inf
        )
        fint[:, max_peaks : max_peaks + 5] = np.where(
            nz, rng.uniform(0.0, 0.05, (m, 5)) * base, 0.0
        )
        # Clip to the fragment range (jitter can leave it).
        oob = (fmz < min_frag_mz) | (fmz > max_frag_mz)
        fmz[oob] = np.inf
        fint[oob] = 0.0
        # Keep the max_peaks most intense.
        cnt = np.isfinite(fmz).sum(axis=1)
        if (cnt > max_peaks).any():
            order = np.argsort(-fint, axis=1, kind="stable")
            rank = np.empty_like(order)
            np.put_along_axis(rank, order, np.arange(width)[None, :], axis=1)
            drop = rank >= max_peaks
            fmz[drop] = np.inf
            fint[drop] = 0.0
        # Guarantee at least one peak.
        empty = ~np.isfinite(fmz).any(axis=1)
        if empty.any():
            fmz[empty, 0] = 0.5 * (min_frag_mz + max_frag_mz)
            fint[empty, 0] = 1.0
        # Sort by m/z, normalise.
        order = np.argsort(fmz, axis=1, kind="stable")
        fmz = np.take_along_axis(fmz, order, axis=1).astype(np.float32)
        fint = np.take_along_axis(fint, order, axis=1).astype(np.float32)
        valid = np.isfinite(fmz)
        fint[~valid] = 0
        norm = np.sqrt((fint.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
        fint = (fint / norm).astype(np.float32)
        mz_out.append(fmz[valid])
        int_out.append(fint[valid])
        cnt_out.append(valid.sum(axis=1))
    counts = np.concatenate(cnt_out).astype(np.int64)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    pmz = t_pmz[template] * (1.0 + rng.uniform(-5e-6, 5e-6, n))
    rt = (t_rt[template] + rng.normal(0.0, 10.0, n)).astype(np.float32)
    out = SpectrumSet(
        np.concatenate(mz_out),
        np.concatenate(int_out),
        indptr,
        pmz.astype(np.float64),
        t_charge[template].astype(np.int32),
        rt,
        template,
    )
    if shuffle:
        out = out.take(rng.permutation(n))
    return out


# --------------------------------------------------------------------------- one large data set in bucket-aligned chunks
BUCKET_WIDTH = 1.0005079  # falcon's precursor bucket rule: round((mz - 1.00794) * z / 1.0005079) // mz_interval


def chunk_mass_ranges(n_chunks: int, mass_range: tuple[float, float] = (700.0, 3500.0)):
    """Neutral-mass slices of ``mass_range`` whose edges sit in the middle between two precursor bucket
    keys, pulled in by 0.03 Da (precursor jitter of 5 ppm and the proton-mass offset of the rule stay
    below that): every precursor bucket of the union of the chunks lies inside ONE chunk, so chunks --
    and any partition of the chunks over GPUs -- are sets of whole buckets."""
    k_lo = int(np.ceil(mass_range[0] / BUCKET_WIDTH))
    k_hi = int(np.floor(mass_range[1] / BUCKET_WIDTH)) - 1
    if k_hi - k_lo < n_chunks:
        raise ValueError("more chunks than precursor buckets in the mass range")
    edges = k_lo + (np.arange(n_chunks + 1) * (k_hi - k_lo)) // n_chunks
    return [((a - 0.5) * BUCKET_WIDTH + 0.03, (b - 0.5) * BUCKET_WIDTH - 0.03) for a, b in zip(edges[:-1], edges[1:])]


def _generate_chunk(args):
    n, seed, rng_ = args
    return generate(n, seed, mass_range=rng_)


def generate_chunks(total: int, n_chunks: int, chunk_ids, seed: int = 1000, workers: int = 1,
                    mass_range: tuple[float, float] = (700.0, 3500.0)) -> SpectrumSet:
    """Chunks ``chunk_ids`` of the ``total``-spectrum data set made of ``n_chunks`` bucket-aligned chunks
    (``chunk_mass_ranges``; chunk c = ``generate(total // n_chunks, seed + c)`` on its mass slice).  The
    data set is the same however the chunks are dealt out; ``workers`` > 1 generates them in worker
    processes (spawned: safe next to an initialised CUDA context)."""
    ranges = chunk_mass_ranges(n_chunks, mass_range)
    jobs = [(total // n_chunks, seed + int(c), ranges[int(c)]) for c in chunk_ids]
    if workers > 1 and len(jobs) > 1:
        import concurrent.futures as cf
        import multiprocessing as mp

        with cf.ProcessPoolExecutor(min(workers, len(jobs)), mp_context=mp.get_context("spawn")) as ex:
            parts = list(ex.map(_generate_chunk, jobs))
    else:
        parts = [_generate_chunk(j) for j in jobs]
    return concat(parts)
