"""MurmurHash3_x86_32 of int32 keys and falcon's ``hash_lookup`` (oracle).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Restates the public MurmurHash3_x86_32 algorithm for the single case falcon
uses: the key is one C ``int`` (4 little-endian bytes, one body block, no
tail, ``len = 4``), exactly what ``sklearn.utils.murmurhash3_32(int, seed,
positive=True)`` hashes (sklearn/utils/murmurhash.pyx:26-30 ->
sklearn/utils/src/MurmurHash3.cpp:105-157).  Published falcon built
``hash_lookup[i] = murmurhash3_32(i, 0, True) % low_dim`` for every mass bin
``i`` (prose /root/reference/README.md:124-131).
"""
from __future__ import annotations

import numpy as np

_C1 = np.uint64(0xCC9E2D51)
_C2 = np.uint64(0x1B873593)
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def murmurhash3_32(keys, seed: int = 0) -> np.ndarray:
    """Unsigned MurmurHash3_x86_32 of int32 ``keys`` (array or scalar)."""
    k = np.atleast_1d(np.asarray(keys)).astype(np.int64).astype(np.uint64) & _M32
    k = (k * _C1) & _M32
    k = _rotl(k, 15)
    k = (k * _C2) & _M32
    h = (np.uint64(seed) & _M32) ^ k
    h = _rotl(h, 13)
    h = (h * np.uint64(5) + np.uint64(0xE6546B64)) & _M32
    h ^= np.uint64(4)  # len
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


def hash_lookup(vec_len: int, low_dim: int, seed: int = 0) -> np.ndarray:
    """``uint32[vec_len]``: hashed column of every mass bin (unsigned ``%``)."""
    h = murmurhash3_32(np.arange(vec_len, dtype=np.int64), seed)
    return (h % np.uint32(low_dim)).astype(np.uint32)
