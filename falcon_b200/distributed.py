"""Multi-GPU plumbing: precursor buckets are independent units, so
each rank clusters its own buckets with no data-path communication; the only
collective is the final gather of labels and cluster representatives (with the
cluster counts for the label offsets), the same running-offset rule the
reference applies per charge (/root/reference/falcon/falcon.py:189-193).

Works with any ``torch.distributed`` backend: NCCL over NVLink on the B200
box, gloo in the CPU tests.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch
import torch.distributed as dist


# Largest bucket that is clustered as ONE unit.  The published sizing rule gives a bucket of >= 10^6 vectors
# 65 536 inverted lists; the device trainer is built for buckets up to that point, and larger
# buckets are cut into pieces with a precursor-tolerance halo (plan_units), each piece indexed on its own.
MAX_BUCKET_ROWS = 1 << 19


def unit_cost(n_queries, n_candidates, exhaustive: bool = True, n_probe: int = 32) -> np.ndarray:
    """Work of clustering ``n_queries`` rows against a bucket of ``n_candidates`` rows: the scan,
    ``q * c`` (exhaustive) or ``q * c * nprobe / nlist`` (IVF sizing rule of SURVEY A.2), plus the
    linear stages."""
    q = np.asarray(n_queries, np.float64)
    c = np.asarray(n_candidates, np.float64)
    cost = q * c
    if not exhaustive:
        nlist = np.maximum(1.0, 2.0 ** np.floor(np.log2(np.maximum(c, 39.0) / 39.0)))
        nlist = np.where(c < 100, 1.0, nlist)
        cost = cost * np.maximum(1.0, np.minimum(np.ceil(nlist / 8), n_probe)) / nlist
    return cost + q


def lpt_assign(cost, world_size: int) -> np.ndarray:
    """Longest-processing-time bin packing: owner rank of every unit (deterministic)."""
    cost = np.asarray(cost, np.float64)
    owner = np.zeros(cost.shape[0], np.int64)
    load = np.zeros(world_size)
    for b in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += cost[b]
    return owner


def shard_buckets(bucket_sizes, world_size: int, exhaustive: bool = True) -> np.ndarray:
    """Rank of every bucket: longest-processing-time bin packing by the scan
    cost ``n_b^2`` (exhaustive) / ``n_b^2 * nprobe / nlist`` (IVF)."""
    sizes = np.asarray(bucket_sizes, np.float64)
    return lpt_assign(unit_cost(sizes, sizes, exhaustive), world_size)


def gather_labels(labels: torch.Tensor, n_clusters: int, group=None):
    """All ranks -> globally unique labels of every rank's spectra.

    ``labels``: this rank's int32 labels (-1 = noise), any length.  Returns
    ``(all_labels, counts)`` where ``all_labels`` is the concatenation over
    ranks (rank order) with rank r's non-noise labels offset by the number of
    clusters of ranks < r, and ``counts`` the per-rank lengths.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = labels.device
    meta = torch.tensor([labels.shape[0], n_clusters], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_h = metas.view(world, 2).cpu().tolist()  # one synchronisation for all ranks' counts
    lens = [int(m[0]) for m in metas_h]
    ncl = [int(m[1]) for m in metas_h]
    offset = sum(ncl[:rank])
    shifted = torch.where(labels >= 0, labels + offset, labels)
    max_len = max(lens) if lens else 0
    padded = torch.full((max_len,), -1, dtype=labels.dtype, device=dev)
    padded[: labels.shape[0]] = shifted
    out = torch.empty(world * max_len, dtype=labels.dtype, device=dev)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * max_len: r * max_len + lens[r]] for r in range(world)]
    return torch.cat(parts) if parts else out, lens


def gather_labels_padded(labels: torch.Tensor, n_clusters: int, max_len: int, group=None):
    """``gather_labels`` in ONE collective and without a host synchronisation, for callers that know
    an upper bound ``max_len`` of every rank's length (a fixed batch size): each rank sends
    ``[len, n_clusters, labels..., padding]``; the running label offsets
    (/root/reference/falcon/falcon.py:189-193) are computed on the device from the gathered headers.

    Returns ``(padded, lens)``: ``padded[r, :lens[r]]`` are rank r's globally unique labels (-1 = noise
    and beyond ``lens[r]``), ``lens`` an int64 device tensor."""
    world = dist.get_world_size(group)
    dev = labels.device
    n = labels.shape[0]
    if n > max_len:
        raise ValueError(f"{n} labels exceed max_len = {max_len}")
    buf = torch.empty(max_len + 2, dtype=torch.int32, device=dev)
    buf[:2] = torch.tensor([n, n_clusters], dtype=torch.int32).to(dev, non_blocking=True)
    buf[2: 2 + n] = labels
    out = torch.empty(world * (max_len + 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, max_len + 2)
    lens = out[:, 0].to(torch.int64)
    ncl = out[:, 1].to(torch.int64)
    offsets = (torch.cumsum(ncl, 0) - ncl).to(torch.int32)
    lab = out[:, 2:]
    valid = torch.arange(max_len, device=dev)[None, :] < lens[:, None]
    padded = torch.where(valid & (lab >= 0), lab + offsets[:, None], torch.full_like(lab, -1))
    return padded, lens


def gather_representatives(representatives: torch.Tensor, n_spectra: int, group=None) -> torch.Tensor:
    """All ranks -> the representatives (medoid spectrum indices) of every cluster,
    in the global label order of ``gather_labels``: rank r's indices are offset by
    the number of spectra of ranks < r, i.e. they index the concatenated labels."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = representatives.device
    meta = torch.tensor([representatives.shape[0], n_spectra], dtype=torch.int64, device=dev)
    metas = torch.empty(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas_h = metas.view(world, 2).cpu().tolist()
    lens = [int(m[0]) for m in metas_h]
    offset = sum(int(m[1]) for m in metas_h[:rank])
    max_len = max(lens) if lens else 0
    padded = torch.full((max_len,), -1, dtype=torch.int64, device=dev)
    padded[: representatives.shape[0]] = representatives.to(torch.int64) + offset
    out = torch.empty(world * max_len, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * max_len: r * max_len + lens[r]] for r in range(world)]
    return torch.cat(parts) if parts else out


class PeerLabelGather:
    """Label gather over NVLink peer memory, without a collective kernel.

    One symmetric buffer per rank (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped
    into every peer over NVLink / NVSwitch) holds ``depth`` slots of ``4 + max_len`` int32.  ``scatter``
    replaces the single-GPU path's final ``flc_scatter32``: the kernel that puts the labels back into input
    order writes them into this rank's slot (``flc_scatter_labels_peers``, local stores).  On a side stream a
    barrier on the peers' signal pads follows, then ``flc_relabel_gathered`` pulls every peer's slot with
    coalesced loads over NVLink and applies the running label offsets
    (/root/reference/falcon/falcon.py:189-193) from the slot headers.  The transfer is off the compute
    stream, which never waits for a peer within a batch -- no NCCL kernel competing for SMs with the
    persistent kernels of the next batch, no staging copy; the compute stream only waits for the barrier of
    batch k - 1 before it overwrites the slot of batch k - 2 (``depth`` = 2): passing that barrier means
    every peer has finished pulling the older slot.

    Raises if symmetric memory cannot be set up (callers fall back to ``gather_labels_padded``)."""

    def __init__(self, max_len: int, device, group=None, depth: int = 2):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.max_len = (int(max_len) + 3) // 4 * 4
        self.slot, self.depth = self.max_len + 4, depth
        self.device = torch.device(device)
        self.buf = symm.empty(depth * self.slot, dtype=torch.int32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._own = (C.c_void_p * 1)(ptrs[self.rank])
        self._slots = [(C.c_void_p * self.world)(*[p + 4 * b * self.slot for p in ptrs]) for b in range(depth)]
        self.side = torch.cuda.Stream(device=self.device)
        self.k = 0
        self.events = [None] * depth
        self.last = None

    def scatter(self, sorted_labels: torch.Tensor, order, n_clusters) -> torch.Tensor:
        """Labels of this batch in input order (a view of this rank's slot); the gathered, globally unique
        labels of all ranks arrive in ``self.last`` = (padded [world, max_len], lens) on ``self.side``."""
        from ._lib import check, lib, ptr

        n = int(sorted_labels.shape[0])
        if n > self.max_len:
            raise ValueError(f"{n} labels exceed max_len = {self.max_len}")
        b = self.k % self.depth
        main = torch.cuda.current_stream()
        prev = self.events[(self.k - 1) % self.depth]
        if prev is not None:
            main.wait_event(prev)  # every peer has pulled the slot about to be overwritten (class docstring)
        offset = b * self.slot
        nc_dev = n_clusters if isinstance(n_clusters, torch.Tensor) else None
        check(lib.flc_scatter_labels_peers(ptr(sorted_labels), ptr(order), n, ptr(nc_dev),
                                           0 if nc_dev is not None else int(n_clusters), self._own, 1, offset,
                                           main.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.hdl.barrier(channel=0)
            out = torch.empty((self.world, self.max_len), dtype=torch.int32, device=self.device)
            lens = torch.empty(self.world, dtype=torch.int64, device=self.device)
            check(lib.flc_relabel_gathered(self._slots[b], self.world, self.max_len, ptr(out), ptr(lens),
                                           self.side.cuda_stream))
            done = torch.cuda.Event()
            done.record(self.side)
        self.events[b] = done
        self.k += 1
        self.last = (out, lens)
        return self.buf[offset + 4: offset + 4 + n]

    def wait(self):
        """Make the current stream wait for the gathers issued so far."""
        torch.cuda.current_stream().wait_stream(self.side)


# --------------------------------------------------------------------------- one data set over N GPUs
def same_partition(a, b) -> bool:
    """Two label arrays describe the same clustering up to renaming (-1 = noise must coincide)."""
    a, b = np.asarray(a, np.int64), np.asarray(b, np.int64)
    if a.shape != b.shape or ((a < 0) != (b < 0)).any():
        return False
    m = a >= 0
    if not m.any():
        return True
    pa, pb = a[m], b[m]
    # a -> b and b -> a must both be functions
    ua, ia = np.unique(pa, return_index=True)
    ub, ib = np.unique(pb, return_index=True)
    return bool(ua.shape == ub.shape and (pb[ia][np.searchsorted(ua, pa)] == pb).all()
                and (pa[ib][np.searchsorted(ub, pb)] == pa).all())


def plan_units(bucket_ptr, mz_sorted, world_size: int, tol: float, tol_mode: str, exhaustive: bool = True,
               bucket_cap=None, n_probe: int = 32):
    """Work units of one bucket-sorted data set and their owner ranks.

    A precursor bucket is an independent unit (the published pipeline never searches across buckets).
    A bucket of more than ``bucket_cap`` rows is cut into contiguous pieces of its precursor-m/z order; a
    piece owns the queries ``[q0, q1)`` and additionally loads the candidates inside one precursor
    tolerance of its m/z range, ``[c0, c1)`` (the halo), so its sparse rows are complete; DBSCAN of a cut
    bucket runs on one rank after the rows have been gathered (``cluster_sharded``).  ``bucket_cap=None``
    picks the cap that keeps any one unit below a rank's fair share of the scan cost.

    Returns a dict of int64 arrays, one entry per unit, in row order: ``bucket, q0, q1, c0, c1, piece
    (0/1), owner``; rows are positions in the bucket-sorted order."""
    bptr = np.asarray(bucket_ptr, np.int64)
    mz = np.asarray(mz_sorted, np.float64)
    sizes = np.diff(bptr)
    if bucket_cap is None:
        # no unit above a rank's fair share of the scan cost, and none above MAX_BUCKET_ROWS (also on one GPU)
        total = float(unit_cost(sizes, sizes, exhaustive, n_probe).sum())
        # (with the IVF index a cut bucket would train one index per piece and no longer give the single-GPU
        # result, so there only MAX_BUCKET_ROWS cuts)
        bucket_cap = min(MAX_BUCKET_ROWS, max(4096, int(np.sqrt(total / max(world_size, 1))))) \
            if (world_size > 1 and exhaustive) else MAX_BUCKET_ROWS
    bucket_cap = max(int(bucket_cap), 1)
    cols = {k: [] for k in ("bucket", "q0", "q1", "c0", "c1", "piece")}

    def add(b, q0, q1, c0, c1, piece):
        for k, v in zip(cols, (b, q0, q1, c0, c1, piece)):
            cols[k].append(v)

    big = sizes > bucket_cap
    for b in range(sizes.shape[0]):
        s, e = int(bptr[b]), int(bptr[b + 1])
        if not big[b]:
            add(b, s, e, s, e, 0)
            continue
        k = -(-(e - s) // bucket_cap)
        seg = mz[s:e]
        # |dm| < tol (Da) or |dm| / mz_candidate * 1e6 < tol (ppm): a superset of the candidates in reach
        halo = tol if tol_mode == "Da" else tol * 1e-6 * float(seg[-1]) * (1.0 + 1e-6)
        halo = halo * (1.0 + 1e-9) + 1e-12
        for i in range(k):
            q0, q1 = s + (e - s) * i // k, s + (e - s) * (i + 1) // k
            c0 = s + int(np.searchsorted(seg, mz[q0] - halo, "left"))
            c1 = s + int(np.searchsorted(seg, mz[q1 - 1] + halo, "right"))
            add(b, q0, q1, c0, c1, 1)
    units = {k: np.asarray(v, np.int64) for k, v in cols.items()}
    units["owner"] = lpt_assign(unit_cost(units["q1"] - units["q0"], units["c1"] - units["c0"], exhaustive, n_probe),
                                world_size)
    return units


def all_gather_var(t: torch.Tensor, group=None):
    """All ranks' 1-D (or [k, len]) tensors of different lengths along the last axis -> list in rank
    order.  Two collectives (lengths, padded payload) and one host synchronisation."""
    world = dist.get_world_size(group)
    dev = t.device
    n = t.shape[-1]
    lens_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens_t, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    lens = [int(x) for x in lens_t.cpu().tolist()]
    mx = max(lens)
    lead = tuple(t.shape[:-1])
    pad = torch.zeros(lead + (mx,), dtype=t.dtype, device=dev)
    pad[..., :n] = t
    out = torch.empty(world * pad.numel(), dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad.reshape(-1), group=group)
    out = out.view((world,) + lead + (mx,))
    return [out[r][..., : lens[r]] for r in range(world)]


def _segments(starts: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Concatenation of the index ranges [starts[i], starts[i] + counts[i])."""
    total = int(counts.sum().item())
    dev = starts.device
    if total == 0:
        return torch.zeros(0, dtype=torch.int64, device=dev)
    ptr = torch.cumsum(counts, 0) - counts
    seg = torch.repeat_interleave(torch.arange(counts.shape[0], device=dev), counts)
    return starts[seg] + (torch.arange(total, device=dev) - ptr[seg])


def assemble_bucket_rows(rows_all: torch.Tensor, ents_all: torch.Tensor, h_lo: torch.Tensor, h_hi: torch.Tensor):
    """Sparse matrix of the buckets ``[h_lo[j], h_hi[j])`` (row ranges of the global bucket order, ascending) from
    the sparse rows the pieces gathered: ``rows_all`` = int32 [2, R] (global row, entry count) in any order,
    ``ents_all`` = int32 [2, E] (global column, float32 distance bits), a row's entries consecutive and rows'
    entries in the order of ``rows_all``.  Rows outside the given buckets are ignored.  Returns
    ``(dist float32 [nnz], cols int32 [nnz], indptr int64 [n + 1], rows int64 [n])``: the buckets concatenated,
    rows in global order, columns renumbered to positions in that concatenation, ``rows`` = their global rows."""
    dev = rows_all.device
    row_g, cnt = rows_all[0].long(), rows_all[1].long()
    ent_start = torch.cumsum(cnt, 0) - cnt
    h_off = torch.cumsum(h_hi - h_lo, 0) - (h_hi - h_lo)  # position of a bucket in the concatenation
    n_home = int((h_hi - h_lo).sum().item())

    def to_pos(gidx):  # global row -> row of the assembled matrix (-1: not in one of the buckets)
        j = torch.searchsorted(h_lo, gidx, right=True) - 1
        jc = j.clamp(min=0)
        ok = (j >= 0) & (gidx < h_hi[jc])
        return torch.where(ok, gidx - h_lo[jc] + h_off[jc], torch.full_like(gidx, -1))

    pos = to_pos(row_g)
    sel = torch.nonzero(pos >= 0).squeeze(1)
    sel = sel[torch.argsort(pos[sel])]
    if sel.shape[0] != n_home or (n_home and not torch.equal(pos[sel], torch.arange(n_home, device=dev))):
        raise RuntimeError("gathered rows of the cut buckets are incomplete")
    cnt_s = cnt[sel]
    indptr = torch.zeros(n_home + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt_s, 0, out=indptr[1:])
    src = _segments(ent_start[sel], cnt_s)
    cols = to_pos(ents_all[0][src].long()).to(torch.int32)
    return (ents_all[1][src].contiguous().view(torch.float32), cols.contiguous(), indptr,
            _segments(h_lo, h_hi - h_lo))


def cluster_sharded(spectra, settings=None, device=None, group=None, bucket_cap=None):
    """Cluster ONE data set on all ranks of ``group``: every rank holds the same host ``spectra``
    (a ``synth.SpectrumSet``-like object), takes its share of the precursor buckets
    (``plan_units``), runs the hot path on them with no data-path communication, and the labels
    and cluster representatives are gathered at the end (running label offset of
    /root/reference/falcon/falcon.py:189-193).  A bucket larger than ``bucket_cap`` is cut into
    pieces with a halo of one precursor tolerance; the pieces' sparse rows are gathered and the
    bucket's DBSCAN runs on one rank, so the result is the single-GPU partition (in exhaustive mode;
    with the IVF index a cut bucket trains one index per piece instead of one per bucket).  Buckets of
    more than ``MAX_BUCKET_ROWS`` rows are always cut, also on one GPU: this is the entry point for data
    whose buckets exceed what ``HotPath.run`` indexes as one unit.

    Exactness of a cut (exhaustive mode): the reference ranks a query's candidates over the whole bucket,
    keeps ``n_neighbors_ann``, and only then applies the precursor tolerance; a piece ranks the candidates
    inside its halo.  Both agree whenever fewer than ``n_neighbors_ann`` rows of the bucket lie within
    ``eps`` (+ the scan margin) of a query -- the eps-cut matrix only holds such rows -- which is the
    regime falcon's defaults (128 candidates, eps 0.1) are chosen for.

    Returns ``(labels, n_clusters, representatives)`` on every rank: int32 labels in INPUT order
    (-1 = noise), the number of clusters, and -- with ``settings.representatives`` -- the input index of
    every cluster's medoid (label order), else None.  Works without an initialised process group
    (one rank)."""
    from . import pipeline
    from ._lib import check, lib, ptr

    s = settings or pipeline.Settings()
    hp = pipeline.HotPath(s, device)
    dev = hp.device
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    n = len(spectra.precursor_mz)
    use_rt = s.rt_tol is not None
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
    if n == 0:
        return np.zeros(0, np.int32), 0, (np.zeros(0, np.int64) if s.representatives else None)
    # ---- every rank: bucket order of the whole data set (12 bytes per spectrum), the plan
    b = hp.bucket_sort(up(spectra.precursor_mz, np.float64), up(spectra.precursor_charge, np.int32),
                       up(spectra.retention_time, np.float32) if use_rt else None)
    order_h = b.order.cpu().numpy().astype(np.int64)
    units = plan_units(b.bucket_ptr.cpu().numpy(), b.mz.cpu().numpy(), world, s.precursor_tol_mass,
                       s.precursor_tol_mode, s.exhaustive, bucket_cap, s.n_probe)
    mine = np.flatnonzero(units["owner"] == rank)
    whole = mine[units["piece"][mine] == 0]
    pieces = mine[units["piece"][mine] == 1]
    # local rows: whole buckets first (their sparse rows never point at a piece), then the pieces' halo ranges
    lo = np.r_[units["q0"][whole], units["c0"][pieces]]
    hi = np.r_[units["q1"][whole], units["c1"][pieces]]
    sizes = hi - lo
    n_whole = int((units["q1"][whole] - units["q0"][whole]).sum())
    n_local = int(sizes.sum())
    l2g_h = np.concatenate([np.arange(a, c) for a, c in zip(lo, hi)] or [np.zeros(0, np.int64)]).astype(np.int64)
    l2g = up(l2g_h, np.int64)
    mz_all, rt_all = b.mz, b.rt
    g = None
    if n_local:
        sub = spectra.take(order_h[l2g_h])
        bptr_l = np.zeros(sizes.shape[0] + 1, np.int64)
        np.cumsum(sizes, out=bptr_l[1:])
        lb = pipeline.Buckets(None, None, mz_all[l2g], rt_all[l2g] if use_rt else None, up(bptr_l, np.int64),
                              int(sizes.shape[0]))
        v = hp.vectorize(up(sub.mz, np.float32), up(sub.intensity, np.float32), up(sub.indptr, np.int64),
                         want_f32=False, max_peaks=int(np.diff(sub.indptr).max(initial=0)))
        ivf = None if s.exhaustive else hp.build_ivf(v, lb)
        g = hp.knn_graph(v, lb, ivf)
        if v.overflow is not None and int(v.overflow.item()) > 0:
            raise RuntimeError("a spectrum hashed to more distinct columns than the sparse rows hold")

    def stage4(graph, rows_g, n_rows, medoid_rows):
        """DBSCAN + split (+ medoids) of complete rows; rows_g = their positions in the global bucket order.
        medoid_rows(labels, n_clusters) -> row of every cluster's representative."""
        if n_rows == 0:
            z = torch.zeros(0, dtype=torch.int32, device=dev)
            return z, 0, torch.zeros(0, dtype=torch.int64, device=dev)
        db, _ = hp.dbscan(graph, n_rows)
        lab, nc = hp.split(db, mz_all[rows_g], values_sorted=True, rt=rt_all[rows_g] if use_rt else None)
        reps = torch.zeros(0, dtype=torch.int64, device=dev)
        if s.representatives and nc:
            reps = rows_g[medoid_rows(lab, nc).long()]
        return lab, nc, reps

    def vectors_of(rows_h, bptr_h):
        """Vectors + bucket descriptor of the given rows of the global bucket order (host index array)."""
        sub_ = spectra.take(order_h[rows_h])
        rows_d = up(rows_h, np.int64)
        bk = pipeline.Buckets(None, None, mz_all[rows_d], rt_all[rows_d] if use_rt else None, up(bptr_h, np.int64),
                              int(bptr_h.shape[0]) - 1)
        vv = hp.vectorize(up(sub_.mz, np.float32), up(sub_.intensity, np.float32), up(sub_.indptr, np.int64),
                          want_f32=False, max_peaks=int(np.diff(sub_.indptr).max(initial=0)))
        return vv, bk

    def whole_medoids(lab, nc):
        # the published rule reads the uncut matrix (HotPath.representatives_exact); whole buckets are the first
        # len(whole) buckets of the local rows, the pieces behind them carry no label here
        full = torch.cat([lab, torch.full((n_local - n_whole,), -1, dtype=torch.int32, device=dev)])
        wb = dataclasses.replace(lb, bucket_ptr=lb.bucket_ptr[: len(whole) + 1], n_buckets=len(whole))
        return hp.representatives_exact(v, wb, ivf, full, nc)

    # ---- whole buckets: the prefix of the local matrix
    if n_whole:
        nnz_w = int(g.indptr[n_whole].item())
        gw = pipeline.KnnGraph(g.dist[:nnz_w], g.indices[:nnz_w], g.indptr[: n_whole + 1], nnz_w)
    lab_w, nc_w, rep_w = stage4(gw if n_whole else None, l2g[:n_whole], n_whole, whole_medoids)
    rows_out, labs_out, reps_out, nc_local = [l2g[:n_whole]], [lab_w], [rep_w], nc_w
    # ---- cut buckets: gather the pieces' query rows, DBSCAN of a bucket on its home rank
    cut = np.unique(units["bucket"][units["piece"] == 1])
    if cut.size:
        row_parts, ent_parts = [], []
        off = n_whole
        for u in pieces:
            c0, c1, q0, q1 = (int(units[k][u]) for k in ("c0", "c1", "q0", "q1"))
            ls, le = off + (q0 - c0), off + (q1 - c0)
            a, e = int(g.indptr[ls].item()), int(g.indptr[le].item())
            row_parts.append(torch.stack([torch.arange(q0, q1, device=dev, dtype=torch.int32),
                                          (g.indptr[ls + 1: le + 1] - g.indptr[ls:le]).to(torch.int32)]))
            ent_parts.append(torch.stack([l2g[g.indices[a:e].long()].to(torch.int32),
                                          g.dist[a:e].view(torch.int32)]))
            off += c1 - c0
        z2 = torch.zeros((2, 0), dtype=torch.int32, device=dev)
        rows_mine = torch.cat(row_parts, 1) if row_parts else z2
        ents_mine = torch.cat(ent_parts, 1) if ent_parts else z2
        if multi:
            rows_all = torch.cat(all_gather_var(rows_mine, group), 1)
            ents_all = torch.cat(all_gather_var(ents_mine, group), 1)
        else:
            rows_all, ents_all = rows_mine, ents_mine
        home = cut[np.arange(cut.size) % world == rank]  # buckets whose stage 4 runs here
        bp_h = b.bucket_ptr.cpu().numpy()
        if home.size:
            h_lo, h_hi = up(bp_h[home], np.int64), up(bp_h[home + 1], np.int64)
            dist_h, cols_h, indptr_h, rows_h = assemble_bucket_rows(rows_all, ents_all, h_lo, h_hi)
            n_home = int(rows_h.shape[0])
            gh = pipeline.KnnGraph(dist_h, cols_h, indptr_h, int(dist_h.shape[0]))

            def home_medoids(lab, nc):
                # uncut rows rank their candidates over the WHOLE bucket before the tolerance filter, which the
                # pieces' halo ranges cannot reproduce: the home rank builds the uncut matrix of its cut buckets
                # itself (exhaustive: a cut bucket has no bucket-wide index)
                sz = bp_h[home + 1] - bp_h[home]
                hb = np.zeros(home.size + 1, np.int64)
                np.cumsum(sz, out=hb[1:])
                vh, bk = vectors_of(rows_h.cpu().numpy(), hb)
                return hp.representatives_exact(vh, bk, None, lab, nc)

            lab_h, nc_h, rep_h = stage4(gh, rows_h, n_home, home_medoids)
            rows_out.append(rows_h)
            labs_out.append(torch.where(lab_h >= 0, lab_h + nc_local, lab_h))
            reps_out.append(rep_h)
            nc_local += nc_h
    # ---- final gather: (row, label) of every rank, labels made disjoint by the running offset
    pay = torch.stack([torch.cat(rows_out).to(torch.int32), torch.cat(labs_out).to(torch.int32)])
    reps_l = torch.cat(reps_out)
    if multi:
        meta = torch.tensor([nc_local], dtype=torch.int64, device=dev)
        metas = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(metas, meta, group=group)
        parts = all_gather_var(pay, group)
        offs = (torch.cumsum(metas, 0) - metas).tolist()
        parts = [torch.stack([p[0], torch.where(p[1] >= 0, p[1] + int(o), p[1])]) for p, o in zip(parts, offs)]
        pay = torch.cat(parts, 1)
        n_clusters = int(metas.sum().item())
        reps_l = torch.cat(all_gather_var(reps_l, group)) if s.representatives else reps_l
    else:
        n_clusters = nc_local
    if pay.shape[1] != n:
        raise RuntimeError(f"{pay.shape[1]} labelled rows gathered for {n} spectra")
    labels_sorted = torch.empty(n, dtype=torch.int32, device=dev)
    labels_sorted[pay[0].long()] = pay[1]
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib.flc_scatter32(ptr(labels_sorted), ptr(b.order), n, ptr(labels),
                            torch.cuda.current_stream().cuda_stream))
    reps = b.order[reps_l].cpu().numpy().astype(np.int64) if s.representatives else None
    return labels.cpu().numpy(), n_clusters, reps
