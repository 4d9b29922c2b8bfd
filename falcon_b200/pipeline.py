"""Device-resident clustering hot path: buckets -> vectors -> (IVF) -> scan ->
k-NN CSR -> DBSCAN -> precursor split.

Python only orchestrates: every stage is one call through the C ABI
(``include/falcon_b200.h``) on raw device pointers; torch owns the buffers and
the stream.  The stage order is the per-charge loop of published falcon
(the surviving skeleton is /root/reference/falcon/falcon.py:151-203):
``vectorize`` -> ``compute_pairwise_distances`` -> ``generate_clusters``.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr

# bf16 inner products of unit vectors differ from the exact value by at most
# 2^-8 (two roundings of relative size 2^-9 on operands whose product sums are
# bounded by 1); candidates are kept down to eps + SCAN_MARGIN and re-scored.
SCAN_MARGIN = 2.0 ** -7


@dataclasses.dataclass
class Settings:
    """falcon's clustering settings (published flag set; snapshot
    defaults at /root/reference/falcon/config.py:77-183 where they survive)."""

    precursor_tol_mass: float = 20.0
    precursor_tol_mode: str = "ppm"
    rt_tol: Optional[float] = None
    fragment_tol: float = 0.05
    eps: float = 0.1
    mz_interval: int = 1
    low_dim: int = 400
    n_neighbors: int = 64
    n_neighbors_ann: int = 128
    batch_size: int = 2 ** 16  # falcon's flag, accepted for compatibility: the device path has no search batches
    n_probe: int = 32
    min_mz: float = 101.0
    max_mz: float = 1500.0
    min_samples: int = 2
    exhaustive: bool = False  # north_star "n_probe = nlist" mode
    hash_seed: int = 0
    kmeans_iters: int = 10
    scan_impl: int = 0  # 0 = tcgen05, 1 = SIMT verification kernel
    eps_cut: bool = True  # keep only dist <= eps in the CSR (what generate_clusters reads)
    representatives: bool = False  # also pick every cluster's medoid (HotPath.representatives, input indices)
    dense_f32: bool = False  # also materialise the dense float32 rows (no stage needs them: all read the sparse copy)
    speculate: bool = True  # steady state: size every stage by the previous batch's counts (upper bounds) and read the
    #                         true counts back once at the end of the step instead of after every stage (HotPath._finish)


def get_dim(min_mz: float, max_mz: float, bin_size: float):
    """(vec_len, min_mz, max_mz) -- /root/reference/falcon/cluster/spectrum.py:172-199."""
    n, s, e = C.c_uint32(), C.c_float(), C.c_float()
    check(lib.flc_get_dim(min_mz, max_mz, bin_size, C.byref(n), C.byref(s), C.byref(e)))
    return int(n.value), float(s.value), float(e.value)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("falcon_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
        if not isinstance(device, torch.device) else device
    check(lib.flc_check_device(dev.index if dev.index is not None else torch.cuda.current_device()))
    return dev


class _Timer:
    """Per-stage CUDA-event timing on the current stream (optional)."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.events: list[tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    def __call__(self, name: str):
        return _TimerCtx(self, name)

    def result(self) -> dict:
        if not self.enabled:
            return {}
        torch.cuda.synchronize()
        out: dict[str, float] = {}
        for name, a, b in self.events:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


class _TimerCtx:
    def __init__(self, timer: _Timer, name: str):
        self.t, self.name = timer, name

    def __enter__(self):
        if self.t.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.t.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.t.events.append((self.name, self.a, b))


@dataclasses.dataclass
class Vectors:
    """Hashed vectors of the spectra (rows in bucket order when ``order`` was given)."""

    x: Optional[torch.Tensor]  # float32 [n, low_dim]
    xb: Optional[torch.Tensor]  # bfloat16 [n, ld_bf16] (tcgen05 scan operand)
    hash_idx: Optional[torch.Tensor]  # int32 [n_peaks]
    ell_idx: Optional[torch.Tensor]  # int16 storage of uint16 [n, ell_width]
    ell_val: Optional[torch.Tensor]  # float32 [n, ell_width]
    ell_nnz: Optional[torch.Tensor]  # int16 storage of uint16 [n]: populated slots per row
    ell_width: int
    n: int
    low_dim: int
    overflow: Optional[torch.Tensor] = None  # int32 [1]: widest row if it exceeded ell_width (else 0)

    def __iter__(self):  # x, xb, hash_idx = hp.vectorize(...)
        return iter((self.x, self.xb, self.hash_idx))


@dataclasses.dataclass
class KnnGraph:
    """Sparse k-NN matrix on the device, rows in bucket order."""

    dist: torch.Tensor  # float32 [nnz]
    indices: torch.Tensor  # int32 [nnz]
    indptr: torch.Tensor  # int64 [n + 1]
    nnz: int
    pair_count: object = 0  # candidate pairs the scan produced (int, or the device-side counter)
    pair_capacity: int = 0  # size of the candidate buffer the scan wrote into (sync=False: to be checked by the caller)

    @property
    def n_pairs(self) -> int:
        if isinstance(self.pair_count, torch.Tensor):
            self.pair_count = int(self.pair_count.item())
        return int(self.pair_count)


@dataclasses.dataclass
class Buckets:
    order: torch.Tensor  # int32 [n] sorted position -> input row
    key: torch.Tensor  # uint32 (as int32 storage) [n]
    mz: torch.Tensor  # float64 [n] precursor m/z in bucket order
    rt: Optional[torch.Tensor]  # float32 [n] in bucket order
    bucket_ptr: torch.Tensor  # int64 [n_buckets + 1] (view of an [n + 1] buffer)
    n_buckets: int  # the true count, or a host-side upper bound (trailing buckets empty)
    n_buckets_dev: Optional[torch.Tensor] = None  # int64 [1]: the true count on the device
    bucket_ptr_full: Optional[torch.Tensor] = None  # the [n + 1] buffer bucket_ptr is a view of


@dataclasses.dataclass
class IvfIndex:
    nlist: torch.Tensor
    nprobe: torch.Tensor
    centroid_ptr: torch.Tensor
    centroids: torch.Tensor
    list_id: torch.Tensor
    probes: torch.Tensor
    max_nprobe: int
    total_centroids: int
    max_ivf_bucket: int = 0


class HotPath:
    """The stage functions, each a thin wrapper over one C-ABI entry point."""

    def __init__(self, settings: Settings | None = None, device=None, profile: bool = False):
        self.s = settings or Settings()
        if self.s.precursor_tol_mode not in _lib.TOL_MODES:
            raise ValueError("Unknown precursor tolerance mode")
        self.device = require_cuda(device)
        self.vec_len, self.min_mz, self.max_mz = get_dim(self.s.min_mz, self.s.max_mz, self.s.fragment_tol)
        self.timer = _Timer(profile)
        self.ld_bf16 = (self.s.low_dim + 7) // 8 * 8

    # ------------------------------------------------------------------ helpers
    def _empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def _ws(self, nbytes: int, name: Optional[str] = None) -> torch.Tensor:
        """Workspace of at least ``nbytes``.  Named workspaces are pure scratch of one
        stage call: they are kept and reused by later calls (grown when needed), so a
        steady-state step does not go through the allocator for them."""
        nbytes = int(nbytes) + 256
        if name is None:
            return torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        cache = self.__dict__.setdefault("_scratch", {})
        buf = cache.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = cache[name] = torch.empty(nbytes + nbytes // 8, dtype=torch.uint8, device=self.device)
        return buf

    # ------------------------------------------------------------------ a5
    def bucket_sort(self, precursor_mz: torch.Tensor, charge: torch.Tensor,
                    rt: Optional[torch.Tensor] = None, n_buckets_cap: Optional[int] = None) -> Buckets:
        """Bucket order of the spectra.  With ``n_buckets_cap`` (a host-side upper bound of the bucket count)
        nothing is read back: ``bucket_ptr`` is padded with empty buckets up to the cap, ``n_buckets`` is the
        cap, and the true count stays on the device in ``Buckets.n_buckets_dev`` (the caller must check it
        against the cap when it next synchronises)."""
        n = precursor_mz.shape[0]
        order = self._empty(n, torch.int32)
        key = self._empty(n, torch.int32)
        mz_sorted = self._empty(n, torch.float64)
        bucket_ptr = self._empty(n + 1, torch.int64)
        nb = C.c_int64(0)
        nb_dev = self._empty(1, torch.int64)
        cap = None if n_buckets_cap is None else max(0, min(int(n_buckets_cap), n))
        with self.timer("bucket_sort"):
            ws = self._ws(lib.flc_bucket_sort_workspace_bytes(n), "bucket_sort")
            check(lib.flc_bucket_sort(ptr(precursor_mz), ptr(charge), n, self.s.mz_interval,
                                      ptr(order), ptr(key), ptr(mz_sorted), ptr(bucket_ptr),
                                      C.byref(nb) if cap is None else None, 0 if cap is None else cap, ptr(nb_dev),
                                      ptr(ws), ws.numel(), _stream()))
            rt_sorted = None
            if rt is not None:
                rt_sorted = self._empty(n, torch.float32)
                check(lib.flc_gather(ptr(rt), ptr(order), n, 4, ptr(rt_sorted), _stream()))
        count = int(nb.value) if cap is None else cap
        return Buckets(order, key, mz_sorted, rt_sorted, bucket_ptr[: count + 1], count, nb_dev, bucket_ptr)

    # ------------------------------------------------------------------ 8f-1: preprocessing
    def preprocess(self, mz: torch.Tensor, intensity: torch.Tensor, indptr: torch.Tensor,
                   precursor_mz: torch.Tensor, charge: Optional[torch.Tensor], min_peaks: int = 5,
                   min_mz_range: float = 250.0, mz_min: Optional[float] = 101.0, mz_max: Optional[float] = 1500.0,
                   remove_precursor_tolerance: Optional[float] = 1.5, min_intensity: Optional[float] = 0.01,
                   max_peaks_used: Optional[int] = 50, scaling: Optional[str] = "off"):
        """``process_spectrum`` (/root/reference/falcon/cluster/spectrum.py:73-169) for all spectra
        at once; defaults are falcon's (/root/reference/falcon/config.py:127-183).  Returns
        (out_mz, out_intensity, out_indptr, valid): compact CSR peaks (rejected spectra keep none)
        and the uint8 validity mask."""
        if scaling == "off":
            scaling = None
        if scaling not in _lib.SCALING:
            raise ValueError("Unknown intensity scaling")
        n = indptr.shape[0] - 1
        n_peaks = int(mz.shape[0])
        out_mz = self._empty(max(n_peaks, 1), torch.float32)
        out_int = self._empty(max(n_peaks, 1), torch.float32)
        out_indptr = self._empty(n + 1, torch.int64)
        valid = self._empty(max(n, 1), torch.uint8)
        total = C.c_int64(0)
        nan = float("nan")
        with self.timer("preprocess"):
            ws = self._ws(lib.flc_preprocess_workspace_bytes(n, n_peaks), "preprocess")
            check(lib.flc_preprocess(ptr(mz), ptr(intensity), ptr(indptr), n, n_peaks, ptr(precursor_mz), ptr(charge),
                                     int(min_peaks), float(min_mz_range), nan if mz_min is None else float(mz_min),
                                     nan if mz_max is None else float(mz_max),
                                     -1.0 if remove_precursor_tolerance is None else float(remove_precursor_tolerance),
                                     -1.0 if min_intensity is None else float(min_intensity),
                                     0 if max_peaks_used is None else int(max_peaks_used), _lib.SCALING[scaling],
                                     ptr(out_mz), ptr(out_int), ptr(out_indptr), ptr(valid), C.byref(total),
                                     ptr(ws), ws.numel(), _stream()))
        return out_mz[: total.value], out_int[: total.value], out_indptr, valid[:n]

    # ------------------------------------------------------------------ a2-a4
    def _ell_width(self, indptr: torch.Tensor, max_peaks: Optional[int]) -> int:
        if max_peaks is None:  # a row has at most as many non-zeros as the spectrum has peaks
            max_peaks = int((indptr[1:] - indptr[:-1]).max().item()) if indptr.shape[0] > 1 else 0
        return max(8, (min(max_peaks, self.s.low_dim) + 7) // 8 * 8)

    def alloc_vectors(self, n: int, width: int, want_bf16=True, want_f32=False, want_ell=True) -> Vectors:
        d = self.s.low_dim
        x = torch.empty((n, d), dtype=torch.float32, device=self.device) if want_f32 else None
        xb = torch.empty((n, self.ld_bf16), dtype=torch.bfloat16, device=self.device) if want_bf16 else None
        ell_idx = ell_val = ell_nnz = None
        if want_ell and n > 0:
            ell_idx = torch.empty((n, width), dtype=torch.int16, device=self.device)
            ell_val = torch.empty((n, width), dtype=torch.float32, device=self.device)
            ell_nnz = self._empty(n, torch.int16)
        return Vectors(x, xb, None, ell_idx, ell_val, ell_nnz, width if ell_idx is not None else 0, n, d)

    def vectorize_into(self, v: Vectors, mz: torch.Tensor, intensity: torch.Tensor, indptr: torch.Tensor,
                       n: int, order: Optional[torch.Tensor] = None, dest: Optional[torch.Tensor] = None,
                       overflow: Optional[torch.Tensor] = None, hash_idx: Optional[torch.Tensor] = None,
                       norm: bool = True) -> None:
        """Vectorise ``n`` spectra (``indptr`` has n + 1 absolute peak offsets) into the
        rows of ``v``: row r <- spectrum order[r], or spectrum r -> row dest[r]."""
        with self.timer("vectorize"):
            check(lib.flc_vectorize(ptr(mz), ptr(intensity), ptr(indptr), ptr(order), ptr(dest), n,
                                    self.min_mz, self.s.fragment_tol, self.vec_len, v.low_dim, self.s.hash_seed,
                                    1 if norm else 0, ptr(v.x), v.low_dim, ptr(v.xb), self.ld_bf16, ptr(hash_idx),
                                    ptr(v.ell_idx), ptr(v.ell_val), ptr(v.ell_nnz), v.ell_width, ptr(overflow),
                                    _stream()))

    def vectorize(self, mz: torch.Tensor, intensity: torch.Tensor, indptr: torch.Tensor,
                  order: Optional[torch.Tensor] = None, want_bf16: bool = True,
                  want_hash_idx: bool = False, norm: bool = True, want_f32: bool = True,
                  want_ell: bool = True, max_peaks: Optional[int] = None) -> Vectors:
        n = indptr.shape[0] - 1
        width = self._ell_width(indptr, max_peaks) if (want_ell and n > 0) else 0
        v = self.alloc_vectors(n, width, want_bf16, want_f32, want_ell)
        v.hash_idx = self._empty(mz.shape[0], torch.int32) if want_hash_idx else None
        overflow = torch.zeros(1, dtype=torch.int32, device=self.device) if v.ell_idx is not None else None
        self.vectorize_into(v, mz, intensity, indptr, n, order=order, overflow=overflow, hash_idx=v.hash_idx,
                            norm=norm)
        v.overflow = overflow
        return v

    def hash_table(self) -> torch.Tensor:
        out = self._empty(self.vec_len, torch.int32)
        check(lib.flc_hash_table(self.vec_len, self.s.low_dim, self.s.hash_seed, ptr(out), _stream()))
        return out

    # ------------------------------------------------------------------ a6
    def ivf_plan(self, buckets: Buckets, caps: Optional[dict] = None):
        """nlist / nprobe / centroid offsets of every bucket and the totals (total centroids, largest nprobe,
        largest IVF bucket).  With ``caps`` (host-side upper bounds of the three totals) nothing is read back:
        the true totals stay in ``cptr[n_buckets : n_buckets + 3]``."""
        nb = buckets.n_buckets
        nlist = self._empty(nb + 1, torch.int32)
        nprobe = self._empty(nb + 1, torch.int32)
        cptr = self._empty(nb + 3, torch.int64)
        total, maxp, maxb = C.c_int64(0), C.c_int32(0), C.c_int64(0)
        if caps is not None:
            check(lib.flc_ivf_plan(ptr(buckets.bucket_ptr), nb, self.s.n_probe, 0, ptr(nlist), ptr(nprobe),
                                   ptr(cptr), None, None, None, _stream()))
            return nlist, nprobe, cptr, int(caps["total"]), int(caps["maxp"]), int(caps["maxb"])
        check(lib.flc_ivf_plan(ptr(buckets.bucket_ptr), nb, self.s.n_probe, 0, ptr(nlist), ptr(nprobe),
                               ptr(cptr), C.byref(total), C.byref(maxp), C.byref(maxb), _stream()))
        return nlist, nprobe, cptr, int(total.value), int(maxp.value), int(maxb.value)

    def build_ivf(self, v: Vectors, buckets: Buckets,
                  centroids: Optional[torch.Tensor] = None, caps: Optional[dict] = None) -> IvfIndex:
        """IVF index of every bucket: trained here (sparse rows needed), or coarse
        assignment against the given ``centroids`` (dense or sparse rows)."""
        n, d = v.n, v.low_dim
        x = v.x
        ld = x.stride(0) if x is not None else d
        with self.timer("ivf_train"):
            nlist, nprobe, cptr, total, maxp, maxb = self.ivf_plan(buckets, caps)
            if maxb >= (1 << 20):
                raise ValueError(f"a precursor bucket holds {maxb} spectra: buckets of 2^20 rows and more are clustered "
                                 "through falcon_b200.distributed.cluster_sharded, which cuts them (also on one GPU)")
            list_id = self._empty(n, torch.int32)
            probes = torch.empty((n, maxp), dtype=torch.int32, device=self.device)
            if centroids is None:
                if v.ell_idx is None:
                    raise ValueError("k-means trains on the sparse rows: vectorize(..., want_ell=True)")
                centroids = torch.empty((max(total, 1), d), dtype=torch.float32, device=self.device)
                ws = self._ws(lib.flc_kmeans_workspace_bytes(n, buckets.n_buckets, total, maxb, v.ell_width, d), "kmeans")
                check(lib.flc_kmeans_train(ptr(v.ell_idx), ptr(v.ell_val), ptr(v.ell_nnz), v.ell_width,
                                           ptr(v.xb), v.xb.stride(0) if v.xb is not None else 0, n, d,
                                           ptr(buckets.bucket_ptr), buckets.n_buckets, ptr(nlist), ptr(cptr),
                                           total, maxb, self.s.kmeans_iters, ptr(centroids), ptr(nprobe), maxp,
                                           ptr(list_id), ptr(probes), ptr(ws), ws.numel(), _stream()))
                return IvfIndex(nlist, nprobe, cptr, centroids, list_id, probes, maxp, total, maxb)
            if centroids.shape[0] < total:
                raise ValueError("centroids array too small for the bucket plan")
        with self.timer("ivf_assign"):
            check(lib.flc_ivf_assign(ptr(x), ld, n, d, ptr(buckets.bucket_ptr), buckets.n_buckets,
                                     ptr(nlist), ptr(nprobe), ptr(cptr), ptr(centroids), maxp,
                                     ptr(v.ell_idx), ptr(v.ell_val), v.ell_width,
                                     ptr(list_id), ptr(probes), _stream()))
        return IvfIndex(nlist, nprobe, cptr, centroids, list_id, probes, maxp, total)

    # ------------------------------------------------------------------ a7-a9
    def scan_threshold(self, eps_cut: Optional[bool] = None) -> float:
        if not (self.s.eps_cut if eps_cut is None else eps_cut):
            return -math.inf
        return float(np.float32(1.0) - np.float32(self.s.eps) - np.float32(SCAN_MARGIN))

    def knn_graph(self, v: Vectors, buckets: Buckets,
                  ivf: Optional[IvfIndex] = None, pair_capacity: Optional[int] = None, sync: bool = True,
                  eps_cut: Optional[bool] = None, query_mask: Optional[torch.Tensor] = None) -> KnnGraph:
        """Sparse k-NN matrix.  ``sync=False``: nothing is read back -- ``dist`` / ``indices`` keep their
        capacity, ``nnz`` is -1 (it is ``indptr[n]``), and the caller must compare ``pair_count`` with
        ``pair_capacity`` when it next synchronises (a scan that overflowed its buffer produced an
        incomplete matrix)."""
        n, d = v.n, v.low_dim
        x, xb = v.x, v.xb
        s = self.s
        if s.rt_tol is not None and buckets.rt is None:
            raise ValueError("rt_tol is set but no retention times were given")
        cut = s.eps_cut if eps_cut is None else eps_cut
        thr = self.scan_threshold(cut)
        if pair_capacity is None:
            pair_capacity = 32 * n + (1 << 20) if cut else None
        if pair_capacity is None:
            sizes = (buckets.bucket_ptr[1:] - buckets.bucket_ptr[:-1])
            pair_capacity = int((sizes * sizes).sum().item()) + 1024
        pair_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        ws = self._ws(lib.flc_scan_workspace_bytes(n, buckets.n_buckets), "scan")
        while True:
            pairs = self._ws(8 * pair_capacity, "pairs")
            with self.timer("scan"):
                check(lib.flc_scan_pairs(ptr(xb), xb.stride(0), n, d, ptr(buckets.bucket_ptr),
                                         buckets.n_buckets,
                                         ptr(ivf.list_id) if ivf else None, ptr(ivf.probes) if ivf else None,
                                         ivf.max_nprobe if ivf else 0, ptr(ivf.nlist) if ivf else None,
                                         thr, s.scan_impl, ptr(pairs), pair_capacity, ptr(pair_count),
                                         ptr(ws), ws.numel(), _stream()))
            # no host round trip here: flc_knn_csr works off the device-side count and reports an
            # overflow of the candidate buffer together with nnz (one synchronisation)
            nnz_cap = max(1, min(pair_capacity, n * s.n_neighbors))
            dist = self._empty(nnz_cap, torch.float32)
            indices = self._empty(nnz_cap, torch.int32)
            indptr = self._empty(n + 1, torch.int64)
            nnz = C.c_int64(0)
            try:
                with self.timer("knn_csr"):
                    ws2 = self._ws(lib.flc_knn_csr_workspace_bytes(n, pair_capacity), "knn_csr")
                    check(lib.flc_knn_csr(ptr(pairs), ptr(pair_count), pair_capacity,
                                          ptr(x), x.stride(0) if x is not None else d,
                                          ptr(v.ell_idx), ptr(v.ell_val), v.ell_width, n, d,
                                          ptr(buckets.mz), ptr(buckets.rt) if s.rt_tol is not None else None,
                                          ptr(ivf.list_id) if ivf else None, ptr(ivf.probes) if ivf else None,
                                          ivf.max_nprobe if ivf else 0,
                                          s.precursor_tol_mass, _lib.TOL_MODES[s.precursor_tol_mode],
                                          -1.0 if s.rt_tol is None else float(s.rt_tol),
                                          s.n_neighbors, s.n_neighbors_ann,
                                          float(np.float32(s.eps)) if cut else float("nan"), ptr(query_mask),
                                          ptr(dist), ptr(indices), nnz_cap, ptr(indptr),
                                          C.byref(nnz) if sync else None, ptr(ws2), ws2.numel(), _stream()))
                if not sync:
                    return KnnGraph(dist, indices, indptr, -1, pair_count, pair_capacity)
                break
            except _lib.CapacityError:
                n_pairs = int(pair_count.item())
                if n_pairs <= pair_capacity:
                    raise
                pair_capacity = n_pairs + 1024  # candidate buffer was too small: rescan
        return KnnGraph(dist[: nnz.value], indices[: nnz.value], indptr, int(nnz.value), pair_count)

    # ------------------------------------------------------------------ a10
    def dbscan(self, g: KnnGraph, n: int, n_sweeps: int = 0):
        """DBSCAN labels and the number of clusters.  ``n_sweeps`` > 0: exactly that many propagation sweeps
        and nothing read back -- returns (labels, n_clusters int64 device tensor, unsettled int32 device
        tensor: non-zero if the sweeps were not enough)."""
        labels = self._empty(n, torch.int32)
        nc, used = C.c_int64(0), C.c_int32(0)
        sync = n_sweeps <= 0
        nc_dev = None if sync else self._empty(1, torch.int64)
        unsettled = None if sync else self._empty(1, torch.int32)
        with self.timer("dbscan"):
            ws = self._ws(lib.flc_dbscan_workspace_bytes(n), "dbscan")
            check(lib.flc_dbscan(ptr(g.dist), ptr(g.indices), ptr(g.indptr), n,
                                 float(np.float32(self.s.eps)), self.s.min_samples, ptr(labels),
                                 C.byref(nc) if sync else None, ptr(nc_dev), max(0, int(n_sweeps)), C.byref(used),
                                 ptr(unsettled), ptr(ws), ws.numel(), _stream()))
        if sync:
            self.dbscan_sweeps = int(used.value)
            return labels, int(nc.value)
        return labels, nc_dev, unsettled

    # ------------------------------------------------------------------ a11-a15
    def split(self, labels: torch.Tensor, mz: torch.Tensor, values_sorted: bool,
              rt: Optional[torch.Tensor] = None, sync: bool = True):
        """``_postprocess_cluster`` of every DBSCAN cluster (/root/reference/falcon/cluster/cluster.py:362-455);
        with ``Settings.rt_tol`` also the retention-time cut (:418-429), which needs ``rt``."""
        n = labels.shape[0]
        out = self._empty(n, torch.int32)
        nc = C.c_int64(0)
        nc_dev = None if sync else self._empty(1, torch.int64)
        with_rt = self.s.rt_tol is not None
        if with_rt:
            if rt is None:
                raise ValueError("rt_tol is set but no retention times were given")
            rt = rt.to(torch.float64)
        with self.timer("split"):
            ws = self._ws(lib.flc_split_workspace_bytes(n, 1 if with_rt else 0), "split")
            check(lib.flc_split_clusters(ptr(labels), ptr(mz), ptr(rt) if with_rt else None, n,
                                         self.s.precursor_tol_mass,
                                         _lib.TOL_MODES[self.s.precursor_tol_mode],
                                         -1.0 if self.s.rt_tol is None else float(self.s.rt_tol),
                                         self.s.min_samples, 1 if values_sorted else 0, ptr(out),
                                         C.byref(nc) if sync else None, ptr(nc_dev), ptr(ws), ws.numel(), _stream()))
        return out, (int(nc.value) if sync else nc_dev)

    # ------------------------------------------------------------------ a16
    def medoids(self, g: KnnGraph, labels: torch.Tensor, n_clusters: int) -> torch.Tensor:
        """Row index (in the matrix' row order) of the representative of every cluster."""
        out = self._empty(n_clusters, torch.int32)
        with self.timer("medoids"):
            ws = self._ws(lib.flc_medoids_workspace_bytes(n_clusters), "medoids")
            check(lib.flc_medoids(ptr(g.dist), ptr(g.indices), ptr(g.indptr), labels.shape[0], ptr(labels),
                                  n_clusters, ptr(out), ptr(ws), ws.numel(), _stream()))
        return out

    def uncut_graph_ranges(self, v: Vectors, buckets: Buckets, ivf: Optional[IvfIndex], max_pairs: int = 1 << 27,
                           query_mask: Optional[torch.Tensor] = None):
        """The FULL ``n_neighbors`` matrix (no eps cut) one range of buckets at a time: a bucket of b rows yields
        b * b candidate pairs, ``max_pairs`` bounds a range (at least one bucket).  Yields ``(r0, r1, graph)``:
        ``graph`` has n rows, of which rows ``r0 .. r1`` (the range's buckets) are populated -- only those with a
        non-zero byte in ``query_mask`` (uint8 [n]) when one is given."""
        bptr = buckets.bucket_ptr.cpu().numpy()
        sizes = np.diff(bptr).astype(np.float64)
        b0 = 0
        while b0 < sizes.shape[0]:
            b1, acc = b0, 0.0
            while b1 < sizes.shape[0] and (b1 == b0 or acc + sizes[b1] ** 2 <= max_pairs):
                acc += sizes[b1] ** 2
                b1 += 1
            sub = dataclasses.replace(buckets, bucket_ptr=buckets.bucket_ptr[b0: b1 + 1], n_buckets=b1 - b0)
            sub_ivf = None if ivf is None else dataclasses.replace(ivf, nlist=ivf.nlist[b0:], nprobe=ivf.nprobe[b0:],
                                                                    centroid_ptr=ivf.centroid_ptr[b0:])
            g = self.knn_graph(v, sub, sub_ivf, pair_capacity=int(acc) + 1024, eps_cut=False, query_mask=query_mask)
            yield int(bptr[b0]), int(bptr[b1]), g
            del g
            b0 = b1

    def representatives_exact(self, v: Vectors, buckets: Buckets, ivf: Optional[IvfIndex], labels: torch.Tensor,
                              n_clusters: int, max_pairs: int = 1 << 27) -> torch.Tensor:
        """Medoid row of every cluster by the published rule: mean distance to the cluster
        members present in the row of the FULL ``n_neighbors`` matrix -- members beyond ``eps`` count as
        well, so the eps-cut matrix the clustering uses is not enough.  The uncut matrix is built for a
        range of buckets at a time (``uncut_graph_ranges``), its medoids taken, and dropped again.
        ``labels``: final labels in row order, numbered in bucket order (what ``split`` returns)."""
        out = torch.full((max(n_clusters, 1),), -1, dtype=torch.int32, device=self.device)[:n_clusters]
        if n_clusters == 0:
            return out
        # clusters of one or two members take their first member whatever the matrix says: only the rows of
        # larger clusters need their uncut neighbour lists
        member = labels >= 0
        size = torch.bincount(labels[member].long(), minlength=n_clusters)
        need = torch.zeros(labels.shape[0], dtype=torch.uint8, device=self.device)
        need[member] = (size[labels[member].long()] > 2).to(torch.uint8)
        for r0, r1, g in self.uncut_graph_ranges(v, buckets, ivf, max_pairs, query_mask=need):
            lab = labels[r0:r1]
            lab = lab[lab >= 0]
            if lab.numel():
                lo, hi = int(lab.min().item()), int(lab.max().item())
                out[lo: hi + 1] = self.medoids(g, labels, n_clusters)[lo: hi + 1]
        return out

    # ------------------------------------------------------------------ whole path
    def _spec_for(self, n: int, keep: bool) -> Optional[dict]:
        """Upper bounds learnt from the previous batch of this size (None: run with a read-back per stage)."""
        if keep or not self.s.speculate or self.s.representatives or not self.s.eps_cut:
            return None
        return self.__dict__.setdefault("_spec", {}).get(n)

    def _learn(self, n: int, nb: int, total: int, maxp: int, maxb: int, pairs: int, sweeps: int) -> None:
        grow = lambda x, lo: int(x + x // 8 + lo)  # noqa: E731 -- 12 % head-room: consecutive batches of a run are alike
        # maxb only decides whether k-means also runs its tiled path: kept exact, and what is checked later is that
        # decision, not the bound.  sweeps: DBSCAN propagation sweeps to launch blind (the last one must change nothing)
        self.__dict__.setdefault("_spec", {})[n] = dict(
            nb=min(n, grow(nb, 64)), total=grow(total, 64), maxp=maxp, maxb=maxb,
            pairs=max(grow(pairs, 1 << 16) + pairs // 2, 1 << 20), sweeps=max(2, sweeps))

    def _cluster_vectors(self, v: Vectors, buckets: Buckets, n: int, keep: bool, spec: Optional[dict] = None):
        if spec is not None:
            return self._cluster_vectors_spec(v, buckets, n, spec)
        ivf = None if self.s.exhaustive else self.build_ivf(v, buckets)
        graph = self.knn_graph(v, buckets, ivf)
        db_labels, _ = self.dbscan(graph, n)
        sorted_labels, n_clusters = self.split(db_labels, buckets.mz, values_sorted=True, rt=buckets.rt)
        if v.overflow is not None and int(v.overflow.item()) > 0:  # the stream was just synchronised
            raise RuntimeError(f"a spectrum hashed to {int(v.overflow.item())} distinct columns but the sparse rows "
                               f"hold {v.ell_width}: pass the true max_peaks")
        self._learn(n, buckets.n_buckets, ivf.total_centroids if ivf else 0, ivf.max_nprobe if ivf else 1,
                    ivf.max_ivf_bucket if ivf else 0, graph.n_pairs, self.dbscan_sweeps + 1)
        sink = getattr(self, "label_sink", None)
        with self.timer("scatter"):
            if sink is not None:  # multi-GPU: the scatter stores into every peer's gather slot (distributed.PeerLabelGather)
                labels = sink.scatter(sorted_labels, buckets.order, n_clusters)
            else:
                labels = self._empty(n, torch.int32)
                check(lib.flc_scatter32(ptr(sorted_labels), ptr(buckets.order), n, ptr(labels), _stream()))
        self.representatives = None
        if self.s.representatives:  # input index of every cluster's medoid
            rows = self.representatives_exact(v, buckets, ivf, sorted_labels, n_clusters)
            self.representatives = self._empty(n_clusters, torch.int32)
            check(lib.flc_gather(ptr(buckets.order), ptr(rows), n_clusters, 4, ptr(self.representatives), _stream()))
        if keep:
            return labels, n_clusters, dict(buckets=buckets, x=v.x, xb=v.xb, vectors=v, ivf=ivf, graph=graph,
                                            db_labels=db_labels, sorted_labels=sorted_labels,
                                            representatives=self.representatives)
        return labels, n_clusters

    def _cluster_vectors_spec(self, v: Vectors, buckets: Buckets, n: int, spec: dict):
        """The stages after vectorisation with NO read-back between them: buffers and grids are sized by
        ``spec`` (upper bounds from the previous batch; the bucket list is padded with empty buckets), every
        count stays on the device, and ONE synchronisation at the end fetches them all.  If a bound turns out
        too small the batch is redone with exact sizes (and the bounds are re-learnt)."""
        ivf = None if self.s.exhaustive else self.build_ivf(v, buckets, caps=spec)
        graph = self.knn_graph(v, buckets, ivf, pair_capacity=spec["pairs"], sync=False)
        db_labels, _, unsettled = self.dbscan(graph, n, n_sweeps=spec["sweeps"])
        sorted_labels, nc_dev = self.split(db_labels, buckets.mz, values_sorted=True, rt=buckets.rt, sync=False)
        nb_cap = buckets.n_buckets
        tail = ivf.centroid_ptr[nb_cap: nb_cap + 3] if ivf is not None else torch.zeros(3, dtype=torch.int64, device=self.device)
        counts = torch.cat([buckets.n_buckets_dev, tail, graph.pair_count.view(torch.int64).view(1), nc_dev,
                            v.overflow.to(torch.int64) if v.overflow is not None else tail[:1] * 0,
                            unsettled.to(torch.int64)]).cpu().tolist()
        nb, total, maxp, maxb, pairs, n_clusters, overflow, unsettled = (int(c) for c in counts)  # the step's one read-back
        if overflow > 0:
            raise RuntimeError(f"a spectrum hashed to {overflow} distinct columns but the sparse rows "
                               f"hold {v.ell_width}: pass the true max_peaks")
        tiled = lambda mb: bool(lib.flc_kmeans_needs_tiled(n, mb, v.ell_width, v.low_dim))  # noqa: E731
        if (nb > nb_cap or total > spec["total"] or maxp > spec["maxp"] or pairs > spec["pairs"] or unsettled
                or (maxb > spec["maxb"] and tiled(maxb) and not tiled(spec["maxb"]))):
            # a bound was too small (or DBSCAN needed more sweeps): this batch again with exact sizes
            exact = dataclasses.replace(buckets, bucket_ptr=buckets.bucket_ptr_full[: nb + 1], n_buckets=nb)
            self.__dict__.get("_spec", {}).pop(n, None)
            self.spec_misses = getattr(self, "spec_misses", 0) + 1
            return self._cluster_vectors(v, exact, n, False, None)
        self._learn(n, nb, total, max(maxp, 1), max(maxb, spec["maxb"]) if not tiled(maxb) else maxb, pairs,
                    spec["sweeps"])
        # the last kernel goes out after the check: with a label sink it also stores into the peers' gather
        # slots, which must happen exactly once per batch
        sink = getattr(self, "label_sink", None)
        with self.timer("scatter"):
            if sink is not None:
                labels = sink.scatter(sorted_labels, buckets.order, n_clusters)
            else:
                labels = self._empty(n, torch.int32)
                check(lib.flc_scatter32(ptr(sorted_labels), ptr(buckets.order), n, ptr(labels), _stream()))
        self.representatives = None
        return labels, n_clusters

    def run(self, mz, intensity, indptr, precursor_mz, charge, rt=None, keep=False, max_peaks=None):
        """All stages on device tensors; returns labels in INPUT order (int32,
        -1 = noise) and the number of clusters.  ``keep`` also returns the
        intermediates (bucket order) for parity checks."""
        n = precursor_mz.shape[0]
        if self.s.rt_tol is not None and rt is None:
            raise ValueError("rt_tol is set but no retention times were given")
        if n == 0:
            empty = self._empty(0, torch.int32)
            return (empty, 0, {}) if keep else (empty, 0)
        spec = self._spec_for(n, keep)
        buckets = self.bucket_sort(precursor_mz, charge, rt if self.s.rt_tol is not None else None,
                                   n_buckets_cap=spec["nb"] if spec else None)
        v = self.vectorize(mz, intensity, indptr, buckets.order, want_f32=keep or self.s.dense_f32,
                           max_peaks=max_peaks)
        return self._cluster_vectors(v, buckets, n, keep, spec)

    def stage_host(self, mz, intensity, indptr, precursor_mz, charge, rt=None,
                   max_peaks: Optional[int] = None, n_chunks: int = 8) -> Optional[dict]:
        """Enqueue the H2D copies of one batch of HOST tensors (pinned memory for real
        overlap) on the copy stream and return at once.  The precursor columns go
        first so that bucketing can start while the peak arrays are still crossing
        PCIe; the peaks follow in chunks, one event each."""
        n = int(precursor_mz.shape[0])
        if self.s.rt_tol is not None and rt is None:
            raise ValueError("rt_tol is set but no retention times were given")
        if n == 0:
            return None
        dev = self.device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = [dict(cap_n=0, cap_p=0, free=None) for _ in range(2)]
            self._staged = 0
        copy = self._copy_stream
        slot = self._slots[self._staged % 2]  # two persistent device-side input slots, used alternately
        if slot.get("pending"):
            raise RuntimeError("stage_host: both input slots hold batches that have not been run yet")
        self._staged += 1
        slot["pending"] = True
        use_rt = rt is not None and self.s.rt_tol is not None
        n_peaks = int(mz.shape[0])
        if slot["cap_n"] < n or slot["cap_p"] < n_peaks or (use_rt and slot.get("rt") is None):
            cap_n, cap_p = max(n, slot["cap_n"]), max(n_peaks, slot["cap_p"])
            torch.cuda.synchronize(dev)  # growing a slot: nothing may still be reading or filling the old buffers
            slot.update(cap_n=cap_n, cap_p=cap_p,
                        pmz=self._empty(cap_n, torch.float64), z=self._empty(cap_n, torch.int32),
                        rt=self._empty(cap_n, torch.float32) if use_rt else None,
                        indptr=self._empty(cap_n + 1, torch.int64),
                        mz=self._empty(cap_p, torch.float32), intensity=self._empty(cap_p, torch.float32))
            copy.wait_stream(torch.cuda.current_stream())
        n_chunks = max(1, min(n_chunks, n))
        bounds = [n * c // n_chunks for c in range(n_chunks + 1)]
        peak_bounds = [int(indptr[b]) for b in bounds]
        if max_peaks is None:
            max_peaks = int((indptr[1:] - indptr[:-1]).max())
        with torch.cuda.stream(copy):
            if slot["free"] is not None:
                copy.wait_event(slot["free"])  # the batch that used this slot last has read its inputs
            pmz_d, z_d, indptr_d = slot["pmz"][:n], slot["z"][:n], slot["indptr"][: n + 1]
            mz_d, in_d = slot["mz"][:n_peaks], slot["intensity"][:n_peaks]
            rt_d = slot["rt"][:n] if use_rt else None
            pmz_d.copy_(precursor_mz, non_blocking=True)
            z_d.copy_(charge, non_blocking=True)
            if use_rt:
                rt_d.copy_(rt, non_blocking=True)
            ev_meta = torch.cuda.Event()
            ev_meta.record(copy)
            indptr_d.copy_(indptr, non_blocking=True)
            events = []
            for c in range(n_chunks):
                a, b = peak_bounds[c], peak_bounds[c + 1]
                mz_d[a:b].copy_(mz[a:b], non_blocking=True)
                in_d[a:b].copy_(intensity[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                events.append(ev)
        return dict(n=n, pmz=pmz_d, z=z_d, rt=rt_d, indptr=indptr_d, mz=mz_d, intensity=in_d, ev_meta=ev_meta,
                    events=events, bounds=bounds, width=self._ell_width(indptr, max_peaks), slot=slot)

    def run_staged(self, st: Optional[dict], labels_out=None):
        """The compute stages of a batch staged with ``stage_host``: every chunk is
        vectorised -- scattered straight into bucket order -- as soon as it has
        landed.  Returns (labels, n_clusters); labels are also copied to
        ``labels_out`` (host) when given."""
        if st is None:
            return self._empty(0, torch.int32), 0
        n = st["n"]
        compute = torch.cuda.current_stream()
        try:
            compute.wait_event(st["ev_meta"])
            spec = self._spec_for(n, False)
            buckets = self.bucket_sort(st["pmz"], st["z"], st["rt"], n_buckets_cap=spec["nb"] if spec else None)
            rank = self._empty(n, torch.int32)  # input position -> bucket-order row
            check(lib.flc_scatter32(None, ptr(buckets.order), n, ptr(rank), _stream()))
            v = self.alloc_vectors(n, st["width"], want_f32=self.s.dense_f32)
            overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
            bounds = st["bounds"]
            for c, ev in enumerate(st["events"]):
                compute.wait_event(ev)
                i0, i1 = bounds[c], bounds[c + 1]
                self.vectorize_into(v, st["mz"], st["intensity"], st["indptr"][i0:], i1 - i0, dest=rank[i0:],
                                    overflow=overflow)
            v.overflow = overflow
        finally:
            # the staged inputs are not read past this point (also when a stage raised: the slot
            # must not stay blocked)
            st["slot"]["free"] = torch.cuda.Event()
            st["slot"]["free"].record(compute)
            st["slot"]["pending"] = False
        labels, n_clusters = self._cluster_vectors(v, buckets, n, False, spec)
        if labels_out is not None:
            labels_out.copy_(labels, non_blocking=True)
        return labels, n_clusters

    def run_host(self, mz, intensity, indptr, precursor_mz, charge, rt=None, labels_out=None,
                 max_peaks: Optional[int] = None, n_chunks: int = 8):
        """The whole path from HOST tensors: ``stage_host`` + ``run_staged``.  Batches
        can be software-pipelined by staging batch i + 1 before running batch i
        (the copies then hide behind the previous batch's kernels)."""
        return self.run_staged(self.stage_host(mz, intensity, indptr, precursor_mz, charge, rt, max_peaks, n_chunks),
                               labels_out)


def cluster_host(spectra, settings: Settings | None = None, device=None, profile=False):
    """End-to-end on HOST arrays (a ``synth.SpectrumSet``-like object): chunked H2D
    copy overlapped with vectorisation, all stages, D2H of the labels.  Returns
    (labels int32 numpy in input order, n_clusters, per-stage ms)."""
    hp = HotPath(settings, device, profile)

    def pin(a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
        return t.pin_memory() if t.numel() else t

    n = len(spectra.precursor_mz)
    out = torch.empty(n, dtype=torch.int32).pin_memory() if n else torch.empty(0, dtype=torch.int32)
    rt = pin(spectra.retention_time, np.float32) if hp.s.rt_tol is not None else None
    _, n_clusters = hp.run_host(pin(spectra.mz, np.float32), pin(spectra.intensity, np.float32),
                                pin(spectra.indptr, np.int64), pin(spectra.precursor_mz, np.float64),
                                pin(spectra.precursor_charge, np.int32), rt, labels_out=out)
    torch.cuda.current_stream().synchronize()
    return out.numpy().copy(), n_clusters, hp.timer.result()
