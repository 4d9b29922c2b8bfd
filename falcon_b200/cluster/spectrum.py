"""Spectrum vectorisation -- mirror of ``falcon.cluster.spectrum``.

Same names and argument meaning as the reference
(/root/reference/falcon/cluster/spectrum.py) and as the published
``to_vector_parallel``; the arithmetic runs in
``flc_vectorize`` on the GPU.
"""
from __future__ import annotations

import collections
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import _lib, pipeline, synth
from .._lib import check, lib, ptr

# /root/reference/falcon/cluster/spectrum.py:13-24
MsmsSpectrumNb = collections.namedtuple(
    "MsmsSpectrumNb",
    ["filename", "identifier", "precursor_mz", "precursor_charge", "retention_time", "mz", "intensity"],
)


def get_dim(min_mz: float, max_mz: float, bin_size: float):
    """Number of bins and true mass range -- spectrum.py:172-199 (float32 arithmetic)."""
    return pipeline.get_dim(min_mz, max_mz, bin_size)


def hash_lookup(vec_len: int, low_dim: int, seed: int = 0) -> np.ndarray:
    """``uint32[vec_len]``: ``murmurhash3_32(i, seed, positive=True) % low_dim`` for every
    mass bin ``i`` -- the table published falcon built with sklearn."""
    dev = pipeline.require_cuda()
    out = torch.empty(int(vec_len), dtype=torch.int32, device=dev)
    check(lib.flc_hash_table(int(vec_len), int(low_dim), int(seed), ptr(out), pipeline._stream()))
    return out.cpu().numpy().view(np.uint32)


def _as_spectrum_set(spectra) -> synth.SpectrumSet:
    if isinstance(spectra, synth.SpectrumSet):
        return spectra
    if len(spectra) and not isinstance(spectra[0], dict):
        spectra = [
            {"mz": s.mz, "intensity": s.intensity, "precursor_mz": s.precursor_mz,
             "precursor_charge": s.precursor_charge, "retention_time": getattr(s, "retention_time", None)}
            for s in spectra
        ]
    return synth.SpectrumSet.from_dicts(list(spectra))


def to_vector_parallel(
    spectra: Union[Sequence[Dict], synth.SpectrumSet],
    dim: int,
    min_mz: float,
    max_mz: float,
    bin_size: float,
    hash_lookup: Optional[np.ndarray] = None,  # noqa: A002 (reference argument name)
    norm: bool = True,
    seed: int = 0,
) -> np.ndarray:
    """Hashed ``float32[n, dim]`` vectors of the spectra.

    ``min_mz``/``max_mz`` are the bounds returned by :func:`get_dim`.  The hash
    is computed in-kernel, so ``hash_lookup`` (kept for signature
    compatibility) must be ``None`` or equal to the MurmurHash3 table.
    """
    if dim <= 0:
        raise ValueError("dim must be positive")
    ss = _as_spectrum_set(spectra)
    n = len(ss)
    if n == 0:
        return np.zeros((0, dim), np.float32)
    dev = pipeline.require_cuda()
    vec_len = int(np.ceil(np.float32(np.float32(max_mz) - np.float32(min_mz)) / np.float32(bin_size)))
    if hash_lookup is not None:
        table = globals()["hash_lookup"]
        expect = table(len(hash_lookup), dim, seed)
        if not np.array_equal(np.asarray(hash_lookup, np.uint32), expect):
            raise NotImplementedError(
                "only the MurmurHash3 feature-hashing table is supported (the hash is computed in-kernel)")
        vec_len = len(hash_lookup)
    mz = torch.from_numpy(ss.mz).to(dev)
    inten = torch.from_numpy(ss.intensity).to(dev)
    indptr = torch.from_numpy(ss.indptr).to(dev)
    out = torch.empty((n, dim), dtype=torch.float32, device=dev)
    check(lib.flc_vectorize(ptr(mz), ptr(inten), ptr(indptr), None, None, n, float(min_mz), float(bin_size),
                            vec_len, dim, seed, 1 if norm else 0, ptr(out), dim, None, 0, None,
                            None, None, None, 0, None, pipeline._stream()))
    return out.cpu().numpy()


def df_row_to_spec(row) -> MsmsSpectrumNb:
    """spectrum.py:299-322."""
    return MsmsSpectrumNb(row["filename"], row["identifier"], row["precursor_mz"], row["precursor_charge"],
                          row["retention_time"], row["mz"], row["intensity"])


def process_spectra(
    spectra: Union[Sequence, synth.SpectrumSet],
    min_peaks: int,
    min_mz_range: float,
    mz_min: Optional[float] = None,
    mz_max: Optional[float] = None,
    remove_precursor_tolerance: Optional[float] = None,
    min_intensity: Optional[float] = None,
    max_peaks_used: Optional[int] = None,
    scaling: Optional[str] = None,
):
    """``process_spectrum`` (spectrum.py:73-169) for a whole batch on the GPU (``flc_preprocess``).

    Returns ``(processed, valid)``: a ``SpectrumSet`` holding only the spectra that pass the
    quality checks (processed peaks, metadata carried over) and the boolean mask over the input."""
    ss = _as_spectrum_set(spectra)
    hp = pipeline.HotPath(pipeline.Settings())
    dev = hp.device
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    mz, inten, indptr, valid = hp.preprocess(
        up(ss.mz, np.float32), up(ss.intensity, np.float32), up(ss.indptr, np.int64),
        up(ss.precursor_mz, np.float64), up(ss.precursor_charge, np.int32), min_peaks, min_mz_range, mz_min, mz_max,
        remove_precursor_tolerance, min_intensity, max_peaks_used, "off" if scaling is None else scaling)
    valid_h = valid.cpu().numpy().astype(bool)
    full = synth.SpectrumSet(mz.cpu().numpy(), inten.cpu().numpy(), indptr.cpu().numpy(), ss.precursor_mz,
                             ss.precursor_charge, ss.retention_time, ss.template)
    return full.take(np.flatnonzero(valid_h)), valid_h


def process_spectrum(
    spectrum,
    min_peaks: int,
    min_mz_range: float,
    mz_min: Optional[float] = None,
    mz_max: Optional[float] = None,
    remove_precursor_tolerance: Optional[float] = None,
    min_intensity: Optional[float] = None,
    max_peaks_used: Optional[int] = None,
    scaling: Optional[str] = None,
) -> Optional[Dict]:
    """Process one spectrum -- same signature and result as the reference (spectrum.py:73-169):
    a spectrum dict (or an object with the ``MsmsSpectrum`` attributes) in, the processed spectrum
    dict or ``None`` out.  Prefer ``process_spectra`` for more than a handful of spectra."""
    get = (lambda k, d=None: spectrum.get(k, d)) if isinstance(spectrum, dict) else \
        (lambda k, d=None: getattr(spectrum, k, d))
    one = {"mz": get("mz"), "intensity": get("intensity"), "precursor_mz": get("precursor_mz"),
           "precursor_charge": get("precursor_charge"), "retention_time": get("retention_time")}
    out, valid = process_spectra([one], min_peaks, min_mz_range, mz_min, mz_max, remove_precursor_tolerance,
                                 min_intensity, max_peaks_used, scaling)
    if not valid[0]:
        return None
    return {"identifier": get("identifier"), "precursor_mz": get("precursor_mz"),
            "precursor_charge": get("precursor_charge"), "mz": out.mz, "intensity": out.intensity,
            "retention_time": get("retention_time"), "filename": get("filename")}
