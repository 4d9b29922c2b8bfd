"""Multi-GPU plumbing: precursor buckets are independent units, so
each rank clusters its own buckets with no data-path communication; the only
collective is the final gather of labels and cluster representatives (with the
cluster counts for the label offsets), the same running-offset rule the
reference applies per charge (/root/reference/falcon/falcon.py:189-193).

Works with any ``torch.distributed`` backend: NCCL over NVLink on the B200
box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_buckets(bucket_sizes, world_size: int, exhaustive: bool = True) -> np.ndarray:
    """Rank of every bucket: longest-processing-time bin packing by the scan
    cost ``n_b^2`` (exhaustive) / ``n_b^2 * nprobe / nlist`` (IVF)."""
    sizes = np.asarray(bucket_sizes, np.float64)
    cost = sizes * sizes
    if not exhaustive:
        nlist = np.maximum(1.0, 2.0 ** np.floor(np.log2(np.maximum(sizes, 39.0) / 39.0)))
        nlist[sizes < 100] = 1.0
        cost = cost * np.maximum(1.0, np.minimum(np.ceil(nlist / 8), 32)) / nlist
    cost = cost + sizes  # linear stages
    owner = np.zeros(sizes.shape[0], np.int64)
    load = np.zeros(world_size)
    for b in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += cost[b]
    return owner


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
