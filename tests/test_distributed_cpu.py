"""world_size-2 gloo test of the label gather and the bucket sharding rule."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from falcon_b200 import distributed as fd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    labels = [torch.tensor([0, -1, 1, 1, 0], dtype=torch.int32), torch.tensor([-1, 0, 0], dtype=torch.int32)][rank]
    ncl = [2, 1][rank]
    out, lens = fd.gather_labels(labels, ncl)
    reps = [torch.tensor([0, 2], dtype=torch.int32), torch.tensor([1], dtype=torch.int32)][rank]
    rep_all = fd.gather_representatives(reps, labels.shape[0])
    padded, plens = fd.gather_labels_padded(labels, ncl, max_len=6)
    # a rank without spectra (an empty shard) still takes part
    empty = labels if rank == 0 else labels[:0]
    padded2, plens2 = fd.gather_labels_padded(empty, ncl if rank == 0 else 0, max_len=5)
    # variable-length gather of [k, len] payloads (cluster_sharded's row / entry exchange), empty part included
    var = fd.all_gather_var(torch.arange(6, dtype=torch.int32).view(2, 3) + 10 * rank if rank == 0
                            else torch.zeros((2, 0), dtype=torch.int32))
    var1 = fd.all_gather_var(torch.arange(rank + 1, dtype=torch.int64))
    q.put((rank, out.tolist(), lens, rep_all.tolist(), padded.tolist(), plens.tolist(), padded2.tolist(),
           plens2.tolist(), [v.tolist() for v in var], [v.tolist() for v in var1]))
    dist.destroy_process_group()


def test_gather_labels_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, out, lens, reps, padded, plens, padded2, plens2, var, var1 in res:
        assert var == [[[0, 1, 2], [3, 4, 5]], [[], []]] and var1 == [[0], [0, 1]]
        assert plens2 == [5, 0] and padded2 == [[0, -1, 1, 1, 0], [-1] * 5]
        assert out == [0, -1, 1, 1, 0, -1, 2, 2] and lens == [5, 3]
        # the single-collective variant: same labels, padded to max_len with -1
        assert plens == [5, 3]
        assert padded == [[0, -1, 1, 1, 0, -1], [-1, 2, 2, -1, -1, -1]]
        # cluster c's representative carries label c in the gathered labels
        assert reps == [0, 2, 6] and [out[i] for i in reps] == [0, 1, 2]


def test_shard_buckets_balances_cost():
    rng = np.random.default_rng(0)
    sizes = rng.integers(10, 5000, 3000)
    for world in (1, 2, 4, 8):
        owner = fd.shard_buckets(sizes, world)
        assert owner.min() >= 0 and owner.max() < world
        cost = sizes.astype(np.float64) ** 2 + sizes
        load = np.bincount(owner, weights=cost, minlength=world)
        assert load.max() / load.mean() < 1.05
    assert (fd.shard_buckets(sizes, 2, exhaustive=False) < 2).all()


def test_sharded_clustering_equals_single_run():
    """buckets are independent, so clustering each rank's buckets separately and merging
    with the running label offset gives the single-run partition.  Oracle on the host, two and three
    'ranks' by the production sharding rule."""
    from oracle import dbscan as odb
    from tests import helpers

    sp = helpers.dataset(2500, 7, 1000.0, 1012.0)
    whole = helpers.oracle_pipeline(sp, exhaustive=True)
    order, bptr = whole["order"], whole["bucket_ptr"]
    sizes = np.diff(bptr)
    for world in (2, 3):
        owner = fd.shard_buckets(sizes, world)
        merged = np.full(len(sp), -1, np.int64)  # labels in the single run's bucket order
        offset = 0
        for r in range(world):
            rows = np.concatenate([np.arange(bptr[b], bptr[b + 1]) for b in np.flatnonzero(owner == r)] or
                                  [np.zeros(0, np.int64)]).astype(np.int64)
            if rows.size == 0:
                continue
            part = helpers.oracle_pipeline(sp.take(order[rows]), exhaustive=True)
            lab = np.asarray(part["labels"], np.int64)
            # the shard is bucket-sorted again by the pipeline: undo its order
            back = np.empty_like(lab)
            back[part["order"]] = lab
            merged[rows] = np.where(back >= 0, back + offset, -1)
            offset += int(lab.max()) + 1 if lab.size and lab.max() >= 0 else 0
        assert odb.same_partition(merged, np.asarray(whole["labels"], np.int64))


def test_plan_units_cuts_oversized_buckets_with_a_tolerance_halo():
    rng = np.random.default_rng(3)
    sizes = np.array([50, 3000, 10, 1200, 700])
    bptr = np.r_[0, np.cumsum(sizes)]
    mz = np.concatenate([np.sort(500.0 + b + rng.uniform(0, 0.5, s)) for b, s in enumerate(sizes)])
    for mode, tol in (("ppm", 20.0), ("Da", 0.01)):
        u = fd.plan_units(bptr, mz, 3, tol, mode, bucket_cap=1000)
        assert (np.bincount(u["bucket"]) == [1, 3, 1, 2, 1]).all()
        # queries: every row exactly once, in order; pieces stay inside their bucket
        assert np.array_equal(np.concatenate([np.arange(a, b) for a, b in zip(u["q0"], u["q1"])]), np.arange(sizes.sum()))
        assert (u["c0"] <= u["q0"]).all() and (u["c1"] >= u["q1"]).all()
        assert (u["c0"] >= bptr[u["bucket"]]).all() and (u["c1"] <= bptr[u["bucket"] + 1]).all()
        assert (u["owner"] >= 0).all() and (u["owner"] < 3).all() and len(set(u["owner"].tolist())) == 3
        for i in np.flatnonzero(u["piece"] == 1):
            q = mz[u["q0"][i]: u["q1"][i]]
            b0, b1 = bptr[u["bucket"][i]], bptr[u["bucket"][i] + 1]
            cand = mz[b0:b1]
            d = np.abs(q[:, None] - cand[None, :])
            reach = (d < tol) if mode == "Da" else (d / cand[None, :] * 1e6 < tol)
            need = np.flatnonzero(reach.any(axis=0)) + b0  # every candidate some query of the piece can reach
            assert need.min() >= u["c0"][i] and need.max() < u["c1"][i]
            assert u["c1"][i] - u["c0"][i] < (b1 - b0)  # and the halo is not the whole bucket
    # a bucket beyond MAX_BUCKET_ROWS is cut whatever the number of ranks
    big = np.sort(700.0 + np.random.default_rng(1).uniform(0, 0.5, fd.MAX_BUCKET_ROWS * 2 + 10))
    u1 = fd.plan_units(np.array([0, big.shape[0]]), big, 1, 20.0, "ppm")
    assert (u1["piece"] == 1).all() and u1["piece"].shape[0] == 3 and (u1["q1"] - u1["q0"]).max() <= fd.MAX_BUCKET_ROWS
    # no cap needed on one rank; the automatic cap leaves small buckets whole
    assert (fd.plan_units(bptr, mz, 1, 20.0, "ppm")["piece"] == 0).all()
    assert (fd.plan_units(bptr, mz, 2, 20.0, "ppm")["piece"] == 0).all()


def test_same_partition():
    assert fd.same_partition([0, 0, 1, -1, 2], [5, 5, 3, -1, 9])
    assert not fd.same_partition([0, 0, 1, -1, 2], [5, 5, 5, -1, 9])
    assert not fd.same_partition([0, 0, 1, -1, 2], [5, 4, 3, -1, 9])
    assert not fd.same_partition([0, 0, 1, -1, 2], [5, 5, 3, 1, 9])
    assert fd.same_partition([-1, -1], [-1, -1]) and not fd.same_partition([0], [0, 0])


def test_bucket_aligned_chunks_hold_whole_buckets():
    """`synth.generate_chunks` (the 10 M / 30 M data sets of bench.py and tools/sweep_30m.py): no precursor bucket
    (charge, interval of the published rule) has spectra in two chunks, whatever subset of chunks a rank takes,
    and the union does not depend on how the chunks are dealt out."""
    from falcon_b200 import synth
    from oracle import ivf as oivf

    total, n_chunks = 48000, 16
    whole = synth.generate_chunks(total, n_chunks, range(n_chunks))
    assert len(whole) == total
    per = total // n_chunks
    keys = oivf.bucket_ids(whole.precursor_mz, whole.precursor_charge)
    owner = {}
    for c in range(n_chunks):
        sl = slice(c * per, (c + 1) * per)
        for k in set(zip(whole.precursor_charge[sl].tolist(), keys[sl].tolist())):
            assert owner.setdefault(k, c) == c, f"bucket {k} in chunks {owner[k]} and {c}"
    # two ranks' shares concatenate to the same data set
    a = synth.generate_chunks(total, n_chunks, range(0, 8))
    b = synth.generate_chunks(total, n_chunks, range(8, 16))
    both = synth.concat([a, b])
    assert np.array_equal(both.mz, whole.mz) and np.array_equal(both.precursor_mz, whole.precursor_mz)
    # chunk edges sit between bucket keys, 0.03 Da away from them
    for lo, hi in synth.chunk_mass_ranges(n_chunks):
        assert abs(((lo - 0.03) / synth.BUCKET_WIDTH) % 1.0 - 0.5) < 1e-9 and hi > lo


def test_assemble_bucket_rows_from_gathered_pieces():
    """The row gather of a cut bucket (cluster_sharded): pieces arrive in rank order, not in row order; rows of
    buckets another rank is home of are ignored; columns are renumbered to the concatenation of the home buckets."""
    # global rows 10..15 = bucket A, 20..23 = bucket B (home here), 30..32 = bucket C (someone else's)
    rows = {10: [(10, .0), (12, .5)], 11: [(11, .0)], 12: [(12, .0), (10, .5), (14, .25)], 13: [], 14: [(14, .0), (12, .25)],
            20: [(20, .0), (22, .125)], 21: [(21, .0)], 22: [(22, .0), (20, .125)],
            30: [(30, .0), (31, .75)], 31: [(31, .0)]}
    order = [14, 20, 30, 21, 10, 11, 31, 12, 13, 22]  # as two "ranks" would send them
    rows_all = torch.tensor([[r for r in order], [len(rows[r]) for r in order]], dtype=torch.int32)
    cols = [c for r in order for c, _ in rows[r]]
    dist = np.float32([d for r in order for _, d in rows[r]])
    ents_all = torch.stack([torch.tensor(cols, dtype=torch.int32), torch.from_numpy(dist.view(np.int32))])
    h_lo, h_hi = torch.tensor([10, 20]), torch.tensor([15, 23])
    d, c, indptr, rg = fd.assemble_bucket_rows(rows_all, ents_all, h_lo, h_hi)
    assert rg.tolist() == [10, 11, 12, 13, 14, 20, 21, 22]
    assert indptr.tolist() == [0, 2, 3, 6, 6, 8, 10, 11, 13]
    pos = {g: i for i, g in enumerate(rg.tolist())}
    want_c = [pos[cc] for r in rg.tolist() for cc, _ in rows[r]]
    want_d = [dd for r in rg.tolist() for _, dd in rows[r]]
    assert c.tolist() == want_c and d.tolist() == want_d
    # a missing row is an error, not a silently shorter matrix
    keep = [i for i, r in enumerate(order) if r != 13]
    sub = rows_all[:, keep]
    with pytest.raises(RuntimeError, match="incomplete"):
        fd.assemble_bucket_rows(sub, ents_all, h_lo, h_hi)
