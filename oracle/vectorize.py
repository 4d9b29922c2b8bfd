"""Spectrum vectorisation oracle: get_dim, binning, hashing, L2 norm.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

import math

import numpy as np

from . import hashing


def get_dim(min_mz: float, max_mz: float, bin_size: float):
    """Number of bins and true m/z bounds.

    Follows /root/reference/falcon/cluster/spectrum.py:172-199.  The reference is
    compiled with the explicit numba signature ``(f4, f4, f4) -> (u4, f4, f4)``
    (:172), so every operation below is a float32 operation -- a float64
    evaluation gives 27981 bins at the defaults instead of 27982.
    """
    f = np.float32
    lo, hi, b = f(min_mz), f(max_mz), f(bin_size)
    start = f(lo - f(np.fmod(lo, b)))
    end = f(f(hi + b) - f(np.fmod(hi, b)))
    n = math.ceil(float(f(f(end - start) / b)))
    return int(n), float(start), float(end)


def bin_indices(mz: np.ndarray, min_mz: float, bin_size: float) -> np.ndarray:
    """Mass-bin index of every peak.

    Follows the expression at /root/reference/falcon/cluster/spectrum.py:291:
    ``floor((mz - min_mz) / bin_size)`` with a float32 ``mz`` and Python-float
    (float64) scalars, i.e. evaluated in float64.
    """
    x = (np.asarray(mz).astype(np.float64) - np.float64(min_mz)) / np.float64(bin_size)
    return np.floor(x).astype(np.int64).astype(np.int32)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit pattern (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def to_vector(
    mz: np.ndarray,
    intensity: np.ndarray,
    indptr: np.ndarray,
    min_mz: float,
    bin_size: float,
    vec_len: int,
    low_dim: int,
    seed: int = 0,
    norm: bool = True,
    return_hash_idx: bool = False,
):
    """Hashed, L2-normalised ``float32[n, low_dim]`` vectors from CSR peaks.

    Published falcon: for every peak, in stored (m/z) order,
    ``v[hash_lookup[bin]] += intensity`` in float32, then ``v /= ||v||``.  It is
    the special case of the snapshot's ``(csr @ transformation).toarray()``
    (/root/reference/falcon/cluster/spectrum.py:240-246) with a 0/1
    ``transformation`` holding a single 1 per row at column
    ``hash_lookup[row]``.

    Conventions fixed here (the GPU kernel follows the same):
      * peaks whose bin falls outside ``[0, vec_len)`` are skipped (the
        reference would index out of bounds);
      * the squared norm is accumulated in float64, the row is multiplied by
        ``1 / sqrt(ss)`` in float64 and rounded once to float32 -- independent of summation
        order, unlike BLAS ``snrm2``; all-zero rows stay zero.
    """
    n = indptr.shape[0] - 1
    bins = bin_indices(mz, min_mz, bin_size)
    table = hashing.hash_lookup(vec_len, low_dim, seed)
    ok = (bins >= 0) & (bins < vec_len)
    hidx = np.full(bins.shape[0], -1, np.int32)
    hidx[ok] = table[bins[ok]].astype(np.int32)
    counts = np.diff(indptr)
    row = np.repeat(np.arange(n, dtype=np.int64), counts)
    out = np.zeros((n, low_dim), np.float32)
    flat = out.reshape(-1)
    # np.add.at applies the float32 additions one by one in array order.
    np.add.at(flat, row[ok] * low_dim + hidx[ok], np.asarray(intensity, np.float32)[ok])
    if norm:
        ss = (out.astype(np.float64) ** 2).sum(axis=1, keepdims=True)
        nz = ss[:, 0] > 0
        o64 = out.astype(np.float64)
        o64[nz] = o64[nz] * (1.0 / np.sqrt(ss[nz]))
        out = o64.astype(np.float32)
    if return_hash_idx:
        return out, hidx
    return out
