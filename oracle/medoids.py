"""Cluster representatives (medoids) from the sparse k-NN matrix.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  **Parity unpinned**: restates
``get_cluster_representatives`` / ``_get_cluster_medoid_index`` of published falcon
0.1.x as recalled in SURVEY.md A.5 (the mounted snapshot only holds the dense
descendant, /root/reference/falcon/cluster/cluster.py:512-553, which needs a
condensed distance matrix the north-star path never builds).
"""
from __future__ import annotations

import numpy as np


def cluster_medoids(data, indices, indptr, labels) -> np.ndarray:
    """Row index of the representative of every cluster ``0 .. labels.max()``.

    Members are visited in ascending row order.  Clusters of at most two members
    return the first member; otherwise the member with the smallest mean distance
    (float32 sum in row order / float32 count) to the cluster members present in
    its sparse row, considered only when more than a quarter of the cluster is
    present; first minimum wins, no eligible row -> first member.  ``-1`` for a
    label nobody carries.
    """
    data = np.asarray(data, np.float32)
    indices = np.asarray(indices)
    indptr = np.asarray(indptr)
    labels = np.asarray(labels)
    n_clusters = int(labels.max()) + 1 if labels.size else 0
    out = np.full(max(n_clusters, 0), -1, np.int32)
    order = np.argsort(labels, kind="stable")
    ls = labels[order]
    for lab in range(n_clusters):
        members = order[np.searchsorted(ls, lab, "left"): np.searchsorted(ls, lab, "right")]
        if members.size == 0:
            continue
        if members.size <= 2:
            out[lab] = members[0]
            continue
        best, best_avg = members[0], np.float32(np.inf)
        for r in members:
            cols = indices[indptr[r]: indptr[r + 1]]
            vals = data[indptr[r]: indptr[r + 1]]
            mask = labels[cols] == lab
            cnt = int(mask.sum())
            if cnt > members.size / 4:
                s = np.float32(0)
                for v in vals[mask]:
                    s = np.float32(s + v)
                avg = np.float32(s / np.float32(cnt))
            else:
                avg = np.float32(np.inf)
            if avg < best_avg:
                best, best_avg = r, avg
        out[lab] = best
    return out
