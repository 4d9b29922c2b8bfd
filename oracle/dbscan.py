"""DBSCAN on the sparse k-NN matrix + precursor-tolerance split (oracle).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

``generate_clusters`` follows published falcon: ``min_samples = 2``
(still at /root/reference/falcon/cluster/cluster.py:66), neighbourhoods
``indices[data <= eps]`` per CSR row, core = ``len(neighbourhood) >= 2``,
sklearn ``dbscan_inner`` (sklearn/cluster/_dbscan_inner.pyx:11-41; how sklearn
forms core samples: sklearn/cluster/_dbscan.py:451-463), then the per-cluster
precursor split that survives in the snapshot
(/root/reference/falcon/cluster/cluster.py:334-509).
"""
from __future__ import annotations

import numpy as np

MIN_SAMPLES = 2


def neighbourhoods(data, indices, indptr, eps):
    """Per-row neighbour ids with ``data <= float32(eps)`` (numpy compares the
    float32 array against the weakly typed Python scalar in float32)."""
    mask = np.asarray(data, np.float32) <= np.float32(eps)
    idx = np.asarray(indices)[mask].astype(np.intp)
    rows = np.repeat(np.arange(indptr.shape[0] - 1), np.diff(indptr))[mask]
    counts = np.bincount(rows, minlength=indptr.shape[0] - 1)
    ptr = np.zeros(indptr.shape[0], np.int64)
    np.cumsum(counts, out=ptr[1:])
    return idx, ptr


def dbscan_sklearn(data, indices, indptr, eps, min_samples: int = MIN_SAMPLES):
    """Labels exactly as sklearn's ``dbscan_inner`` assigns them (-1 = noise)."""
    from sklearn.cluster._dbscan_inner import dbscan_inner

    idx, ptr = neighbourhoods(data, indices, indptr, eps)
    n = indptr.shape[0] - 1
    nb = np.empty(n, dtype=object)
    for i in range(n):
        nb[i] = idx[ptr[i] : ptr[i + 1]]
    core = (np.diff(ptr) >= min_samples).astype(np.uint8)
    labels = np.full(n, -1, dtype=np.intp)
    dbscan_inner(core, nb, labels)
    return labels


def dbscan_min_ancestor(data, indices, indptr, eps, min_samples: int = MIN_SAMPLES):
    """Closed form of ``dbscan_inner``: a point's cluster is
    the rank of the minimum-index core point from which it is reachable along
    edges that leave core points.  Fix-point of min-propagation; this is the
    formulation the CUDA kernel implements."""
    idx, ptr = neighbourhoods(data, indices, indptr, eps)
    n = indptr.shape[0] - 1
    core = np.diff(ptr) >= min_samples
    inf = np.iinfo(np.int64).max
    m = np.where(core, np.arange(n, dtype=np.int64), inf)
    src = np.repeat(np.arange(n), np.diff(ptr))
    e = core[src]
    src, dst = src[e], idx[e]
    while True:
        new = m.copy()
        np.minimum.at(new, dst, m[src])
        if (new == m).all():
            break
        m = new
    seeds = np.flatnonzero(core & (m == np.arange(n)))
    rank = np.full(n + 1, -1, np.int64)
    rank[seeds] = np.arange(seeds.shape[0])
    return np.where(m == inf, -1, rank[np.minimum(m, n)]).astype(np.intp)


# --------------------------------------------------------------------------- precursor split
def _tolerance_distance(lo: float, hi: float, mode) -> float:
    """Complete-linkage distance of two adjacent 1-D groups spanning [lo, hi]
    (/root/reference/falcon/cluster/cluster.py:488-491)."""
    d = hi - lo
    if mode == "ppm":
        d = d / lo * 10**6
    return d


def linkage_1d(values: np.ndarray, mode=None) -> np.ndarray:
    """scipy-style linkage matrix of the 1-D complete-linkage agglomeration.

    Restates /root/reference/falcon/cluster/cluster.py:458-509: groups are
    runs of the sorted values; every step merges the adjacent pair whose
    union has the smallest span (relative to the left group's minimum in ppm
    mode), first minimum wins.
    """
    values = np.asarray(values, np.float64)
    n = values.shape[0]
    order = np.argsort(values)
    lo = [float(values[i]) for i in order]
    hi = list(lo)
    ident = [int(i) for i in order]
    size = [1] * n
    out = np.zeros((max(n - 1, 0), 4), np.float64)
    for step in range(n - 1):
        best, best_i = np.inf, -1
        for i in range(len(lo) - 1):
            d = _tolerance_distance(lo[i], hi[i + 1], mode)
            if d < best:
                best, best_i = d, i
        i = best_i
        out[step] = (ident[i], ident[i + 1], best, size[i] + size[i + 1])
        hi[i], ident[i], size[i] = hi[i + 1], n + step, size[i] + size[i + 1]
        del lo[i + 1], hi[i + 1], ident[i + 1], size[i + 1]
    return out


def split_sorted_1d(sorted_values: np.ndarray, tol: float, mode=None) -> np.ndarray:
    """Flat complete-linkage groups (cut at ``tol``, inclusive) of an ascending
    1-D array, as ids in order of first appearance.  Runs the agglomeration of
    ``linkage_1d`` and stops at the first merge above ``tol`` (complete
    linkage is monotone, so that equals ``fcluster(..., tol, 'distance')``)."""
    v = np.asarray(sorted_values, np.float64)
    starts = list(range(v.shape[0]))
    while len(starts) > 1:
        best, best_i = np.inf, -1
        for i in range(len(starts) - 1):
            end = starts[i + 2] if i + 2 < len(starts) else v.shape[0]
            d = _tolerance_distance(v[starts[i]], v[end - 1], mode)
            if d < best:
                best, best_i = d, i
        if not best <= tol:
            break
        del starts[best_i + 1]
    gid = np.zeros(v.shape[0], np.int64)
    gid[starts] = 1
    return np.cumsum(gid) - 1


def fcluster_ids_by_range_additions(sorted_values: np.ndarray, tol: float, mode=None) -> np.ndarray:
    """0-based ids scipy's ``fcluster(linkage_1d(v), tol, 'distance')`` gives the elements of an ascending 1-D
    array, computed the way the CUDA kernel does (``csrc/dbscan.cu:linkage_ids_one``) instead of by walking the
    dendrogram: ``fcluster`` numbers flat clusters in a depth-first walk from the root -- left subtree, right
    subtree, then the children that are single observations, left before right; a subtree at or below the cut
    gets one id when it is entered (scipy/cluster/_hierarchy.pyx:cluster_monocrit).  Merging nodes A (left) and B
    (right) above the cut therefore shifts the ids of A's flat clusters by off(A) and B's by off(B):
        A, B both single observations: 0, 1     A single only: |B|, 0
        B single only: 0, |A|                   neither: 0, |A|
    (|X| = flat clusters in X), and a flat cluster's id is the sum of the shifts on its path to the root: a
    difference array over the flat clusters (contiguous in sorted order) and one prefix sum.  Checked against
    scipy in tests/test_oracle_golden.py; the kernel is checked against the reference's goldens."""
    v = np.asarray(sorted_values, np.float64)
    m = v.shape[0]
    if m < 2:
        return np.zeros(m, np.int64)
    flat = split_sorted_1d(v, tol, mode)  # flat cluster of every element (ids in sorted order)
    r0 = int(flat[-1]) + 1
    starts = [int(np.searchsorted(flat, k, "left")) for k in range(r0)]  # element offset of every node
    first = list(range(r0))                                               # first flat cluster of every node
    delta = np.zeros(r0 + 1, np.int64)
    while len(starts) > 1:
        best, bi = np.inf, -1
        for i in range(len(starts) - 1):
            end = starts[i + 2] if i + 2 < len(starts) else m
            d = _tolerance_distance(v[starts[i]], v[end - 1], mode)
            if d < best:
                best, bi = d, i
        sa, sb = starts[bi], starts[bi + 1]
        eb = starts[bi + 2] if bi + 2 < len(starts) else m
        fa, fb = first[bi], first[bi + 1]
        fe = first[bi + 2] if bi + 2 < len(first) else r0
        ca, cb = fb - fa, fe - fb
        single_a, single_b = sb - sa == 1, eb - sb == 1
        off_a = (0 if single_b else cb) if single_a else 0
        off_b = (1 if single_a else ca) if single_b else (0 if single_a else ca)
        delta[fa] += off_a
        delta[fb] += off_b - off_a
        delta[fe] -= off_b
        del starts[bi + 1], first[bi + 1]
    return np.cumsum(delta[:r0])[flat]


def postprocess_cluster(mzs, rts, tol, mode, rt_tol, min_samples: int = MIN_SAMPLES):
    """Sub-cluster id (or -1) of every member of ONE DBSCAN cluster and the
    number of sub-clusters kept.  Restates
    /root/reference/falcon/cluster/cluster.py:362-455 (including the
    non-injective ``a * 2 + b * 3`` combination of the m/z and RT cuts,
    :427-429).  Ids are numbered by first appearance among the members."""
    mzs = np.asarray(mzs, np.float64)
    m = mzs.shape[0]
    if m < min_samples:
        return np.full(m, -1, np.int64), 0
    order = np.argsort(mzs, kind="stable")
    assign = np.empty(m, np.int64)
    assign[order] = split_sorted_1d(mzs[order], tol, mode)
    if rt_tol is not None:
        # The reference combines the two cuts as ``a * 2 + b * 3`` (:427-429),
        # which is not injective, so the result depends on the actual ids scipy's
        # ``fcluster`` hands out (a depth-first numbering of the dendrogram).
        # Reproduce it by cutting the restated linkage with scipy itself.
        import scipy.cluster.hierarchy as sch

        a_mz = sch.fcluster(linkage_1d(mzs, mode), tol, "distance").astype(np.int64) - 1
        a_rt = sch.fcluster(linkage_1d(np.asarray(rts, np.float64), None), rt_tol, "distance").astype(np.int64) - 1
        assign = np.unique(a_mz * 2 + a_rt * 3, return_inverse=True)[1]
    _, first, inv, cnt = np.unique(assign, return_index=True, return_inverse=True, return_counts=True)
    keep = cnt >= min_samples
    # number surviving groups by first appearance
    appear = np.argsort(first, kind="stable")
    new_id = np.full(cnt.shape[0], -1, np.int64)
    k = 0
    for g in appear:
        if keep[g]:
            new_id[g] = k
            k += 1
    return new_id[inv], k


def generate_clusters(
    data, indices, indptr, eps, precursor_mzs, rts, tol, mode, rt_tol=None,
    dbscan=dbscan_sklearn,
):
    """Cluster labels (-1 = noise) from the sparse distance matrix."""
    labels = np.asarray(dbscan(data, indices, indptr, eps), np.int64)
    n = labels.shape[0]
    out = np.full(n, -1, np.int64)
    order = np.argsort(labels, kind="stable")
    ls = labels[order]
    start = int(np.searchsorted(ls, 0, "left"))  # skip noise
    nxt = 0
    bounds = np.flatnonzero(np.r_[True, ls[start + 1 :] != ls[start:-1]]) + start if start < n else []
    bounds = list(bounds) + [n]
    pm = np.asarray(precursor_mzs, np.float64)
    rt = None if rts is None else np.asarray(rts)
    for a, b in zip(bounds[:-1], bounds[1:]):
        members = order[a:b]
        sub, k = postprocess_cluster(
            pm[members], None if rt is None else rt[members], tol, mode, rt_tol
        )
        ok = sub >= 0
        out[members[ok]] = sub[ok] + nxt
        nxt += k
    return out


def same_partition(a: np.ndarray, b: np.ndarray) -> bool:
    """Label arrays describe the same clustering up to renaming; -1 (noise)
    must coincide."""
    a, b = np.asarray(a, np.int64), np.asarray(b, np.int64)
    if a.shape != b.shape or ((a < 0) != (b < 0)).any():
        return False
    m = a >= 0
    if not m.any():
        return True
    pa, pb = a[m], b[m]
    fwd = {}
    bwd = {}
    for x, y in zip(pa.tolist(), pb.tolist()):
        if fwd.setdefault(x, y) != y or bwd.setdefault(y, x) != x:
            return False
    return True
