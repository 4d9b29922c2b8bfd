"""Spectrum preprocessing oracle.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows
``process_spectrum`` (/root/reference/falcon/cluster/spectrum.py:73-169, whose
validity check is :27-52 and whose norm is :55-70) step by step.  The
``MsmsSpectrum`` methods it calls live in ``spectrum_utils==0.3.5``
(/root/reference/setup.cfg:36), which is not installed here: **parity unpinned**
for those four methods -- ``set_mz_range`` (inclusive window),
``remove_precursor_peak(tol, 'Da', 0)`` (peaks within tol of
``(neutral_mass / c) + 1.0072766`` for c = z .. 1), ``filter_intensity``
(argsort scan: strictly above ``min_intensity * max``, at most ``max_num_peaks``
most intense; a stable ascending sort decides ties, so later peaks win) and
``scale_intensity`` ('root' = sqrt, 'log' = log2(1 + x), 'rank' = ``max_rank -
argsort(argsort(x)[::-1])``) are restated from the published package.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

PROTON = 1.0072766
SCALING = {None: 0, "root": 1, "log": 2, "rank": 3}


def _valid(mz: np.ndarray, min_peaks: int, min_mz_range: float) -> bool:
    return len(mz) >= min_peaks and len(mz) > 0 and np.float32(mz[-1]) - np.float32(mz[0]) >= np.float32(min_mz_range)


def process_spectrum(mz, intensity, precursor_mz, precursor_charge, min_peaks, min_mz_range, mz_min=None,
                     mz_max=None, remove_precursor_tolerance=None, min_intensity=None, max_peaks_used=None,
                     scaling=None):
    """(mz, intensity) float32 arrays of the processed spectrum, or None if it is rejected."""
    mz = np.asarray(mz, np.float32)
    intensity = np.asarray(intensity, np.float32)
    keep = np.ones(mz.shape[0], bool)
    if mz_min is not None:
        keep &= mz >= np.float32(mz_min)
    if mz_max is not None:
        keep &= mz <= np.float32(mz_max)
    mz, intensity = mz[keep], intensity[keep]
    if not _valid(mz, min_peaks, min_mz_range):
        return None
    if remove_precursor_tolerance is not None:
        z = max(int(precursor_charge) if precursor_charge else 1, 1)
        neutral = (float(precursor_mz) - PROTON) * z
        keep = np.ones(mz.shape[0], bool)
        for c in range(z, 0, -1):
            keep &= np.abs(mz.astype(np.float64) - (neutral / c + PROTON)) > float(np.float32(remove_precursor_tolerance))
        mz, intensity = mz[keep], intensity[keep]
        if not _valid(mz, min_peaks, min_mz_range):
            return None
    if min_intensity is not None or max_peaks_used is not None:
        thr = np.float32(0.0 if min_intensity is None else min_intensity)
        k = len(mz) if max_peaks_used is None else int(max_peaks_used)
        idx = np.argsort(intensity, kind="stable")
        thr = np.float32(thr * intensity[idx[-1]])
        start = len(idx) - 1
        for i, v in enumerate(intensity[idx]):
            if v > thr:
                start = i
                break
        mask = np.zeros(mz.shape[0], bool)
        mask[idx[max(start, len(idx) - k):]] = True
        mz, intensity = mz[mask], intensity[mask]
        if not _valid(mz, min_peaks, min_mz_range):
            return None
    if scaling == "root":
        intensity = np.sqrt(intensity)
    elif scaling == "log":
        intensity = (np.log1p(intensity.astype(np.float64)) / np.log(2.0)).astype(np.float32)
    elif scaling == "rank":
        max_rank = len(mz) if max_peaks_used is None else int(max_peaks_used)
        order_desc = np.argsort(intensity, kind="stable")[::-1]
        intensity = (max_rank - np.argsort(order_desc, kind="stable")).astype(np.float32)
    elif scaling is not None:
        raise ValueError("Unknown intensity scaling")
    nrm = np.float32(np.sqrt(np.sum(intensity.astype(np.float64) ** 2)))
    return mz, (intensity / nrm).astype(np.float32)


def process_spectra(spectra, **kw):
    """Apply ``process_spectrum`` to a ``SpectrumSet``-like object.  Returns
    (valid mask, mz, intensity, indptr) with the CSR arrays holding only valid spectra's peaks
    (rejected spectra keep an empty range)."""
    n = len(spectra.precursor_mz)
    valid = np.zeros(n, bool)
    out_mz, out_int, indptr = [], [], [0]
    for i in range(n):
        a, b = int(spectra.indptr[i]), int(spectra.indptr[i + 1])
        r = process_spectrum(spectra.mz[a:b], spectra.intensity[a:b], spectra.precursor_mz[i],
                             spectra.precursor_charge[i], **kw)
        if r is not None:
            valid[i] = True
            out_mz.append(r[0])
            out_int.append(r[1])
            indptr.append(indptr[-1] + len(r[0]))
        else:
            indptr.append(indptr[-1])
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.float32)  # noqa: E731
    return valid, cat(out_mz), cat(out_int), np.asarray(indptr, np.int64)
