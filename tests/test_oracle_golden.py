"""The oracle against the golden vectors produced by running the reference
(tests/golden/make_golden.py) and against sklearn's own known answers."""
import json
import os

import numpy as np
import pytest

from oracle import dbscan as odb
from oracle import hashing, vectorize


def test_murmur_known_answers(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "murmur.json")))
    keys = np.asarray(g["keys"], np.int32)
    assert hashing.murmurhash3_32(keys, 0).tolist() == g["seed0"]
    assert hashing.murmurhash3_32(keys, 42).tolist() == g["seed42"]
    # sklearn/utils/tests/test_murmurhash.py:11-23
    assert int(hashing.murmurhash3_32(3, 0)[0]) == 847579505
    assert int(hashing.murmurhash3_32(3, 42)[0].astype(np.int32)) == -1823081949
    assert int(hashing.murmurhash3_32(3, 42)[0]) == 2471885347
    # SURVEY B.5
    assert hashing.murmurhash3_32(np.arange(5), 0).astype(np.int32).tolist() == [
        593689054, -68075478, 1085422463, 847579505, 1889779975]


def test_murmur_matches_sklearn_everywhere():
    from sklearn.utils import murmurhash3_32

    keys = np.arange(0, 70000, dtype=np.int32)
    for seed in (0, 1, 1234):
        assert (hashing.murmurhash3_32(keys, seed) == murmurhash3_32(keys, seed, True)).all()
    table = hashing.hash_lookup(27982, 400)
    assert table.dtype == np.uint32 and table.max() < 400
    assert table[3] == 847579505 % 400


def test_get_dim_golden(golden_dir):
    for row in json.load(open(os.path.join(golden_dir, "get_dim.json"))):
        n, s, e = vectorize.get_dim(row["min_mz"], row["max_mz"], row["bin_size"])
        assert (n, s, e) == (row["vec_len"], row["start"], row["end"]), row
    assert vectorize.get_dim(101.0, 1500.0, 0.05) == (27982, 100.95000457763672, 1500.0001220703125)


def test_binning_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "binning.npz"))
    bins = vectorize.bin_indices(g["mz"], float(g["min_mz"]), float(g["bin_size"]))
    assert (bins == g["bins"]).all()
    # the all-float32 evaluation is NOT the reference's
    f32 = np.floor((g["mz"] - np.float32(g["min_mz"])) / np.float32(g["bin_size"])).astype(np.int32)
    assert (f32 != g["bins"]).any()


def test_to_vector_is_csr_times_one_hot(golden_dir):
    """Feature hashing == the snapshot's ``csr @ transformation`` with a 0/1
    transformation (spectrum.py:240-243)."""
    import scipy.sparse as ss

    g = np.load(os.path.join(golden_dir, "binning.npz"))
    vec_len, low_dim = 27982, 400
    n = g["indptr"].shape[0] - 1
    v = vectorize.to_vector(g["mz"], g["intensity"], g["indptr"], float(g["min_mz"]), 0.05,
                            vec_len, low_dim, norm=False)
    table = hashing.hash_lookup(vec_len, low_dim)
    proj = ss.csr_matrix((np.ones(vec_len, np.float32), (np.arange(vec_len), table)), (vec_len, low_dim))
    csr = ss.csr_matrix((g["data"], g["bins"], g["indptr"]), (n, vec_len), np.float32, False)
    ref = (csr @ proj).toarray()
    np.testing.assert_allclose(v, ref, rtol=0, atol=1e-6)
    vn = vectorize.to_vector(g["mz"], g["intensity"], g["indptr"], float(g["min_mz"]), 0.05, vec_len, low_dim)
    np.testing.assert_allclose(np.linalg.norm(vn, axis=1), 1.0, atol=1e-6)


def test_bf16_rounding():
    x = np.float32([1.0, 1.00390625, 1.005859375, 0.1, -2.5, 0.0, 3.3895314e38])
    b = vectorize.to_bf16_bits(x)
    import torch

    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert (b == ref).all()


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    return g


def test_linkage_golden(golden_dir):
    g = _cases(golden_dir)
    for i in range(int(g["n_link"])):
        v, mode, ref = g[f"link_v{i}"], str(g[f"link_mode{i}"]), g[f"link_out{i}"]
        out = odb.linkage_1d(v, mode)
        # With duplicated values the member ids depend on the (unstable) argsort
        # of numba vs numpy; sizes and merge distances do not.
        cols = [0, 1, 3] if np.unique(v).shape[0] == v.shape[0] else [3]
        assert (out[:, cols] == ref[:, cols]).all(), i
        np.testing.assert_allclose(out[:, 2], ref[:, 2], rtol=1e-12)
    # SURVEY B.3
    link = odb.linkage_1d(np.float32([500.000, 500.004, 500.009, 500.030, 500.031, 500.2]), "ppm")
    assert link[:, :2].astype(int).tolist() == [[3, 4], [0, 1], [7, 2], [8, 6], [9, 5]]


def test_postprocess_golden(golden_dir):
    g = _cases(golden_dir)
    n_rt = 0
    for i in range(int(g["n_post"])):
        v, rts, mode = g[f"post_v{i}"], g[f"post_rt{i}"], str(g[f"post_mode{i}"])
        tol, rt_tol = float(g[f"post_tol{i}"]), float(g[f"post_rttol{i}"])
        rt_tol = None if rt_tol < 0 else rt_tol
        n_rt += rt_tol is not None
        ref, k_ref = g[f"post_labels{i}"], int(g[f"post_k{i}"])
        sub, k = odb.postprocess_cluster(v, rts, tol, mode, rt_tol)
        assert k == k_ref, i
        assert odb.same_partition(sub, ref), (i, sub, ref)
    assert n_rt > 10
    sub, k = odb.postprocess_cluster(
        np.float32([500.000, 500.004, 500.009, 500.030, 500.031, 500.2]), None, 20.0, "ppm", None)
    assert k == 2 and sub.tolist() == [0, 0, 0, 1, 1, -1]


def test_split_equals_scipy_fcluster():
    import scipy.cluster.hierarchy as sch

    rng = np.random.default_rng(5)
    for _ in range(200):
        m = int(rng.integers(2, 30))
        mode = "ppm" if rng.random() < 0.5 else None
        v = np.sort(800 + rng.normal(0, 0.02, m))
        if rng.random() < 0.4:
            v[rng.integers(0, m, 3)] = v[m // 2]
            v = np.sort(v)
        tol = 20.0 if mode == "ppm" else 0.016
        ref = sch.fcluster(odb.linkage_1d(v, mode), tol, "distance")
        got = odb.split_sorted_1d(v, tol, mode)
        assert odb.same_partition(got, ref - 1)


def test_group_idx_golden(golden_dir):
    g = _cases(golden_dir)
    assert g["groups_out"].tolist() == [[0, 1], [1, 2], [2, 4], [4, 5], [5, 8]]


def test_fcluster_numbering_as_range_additions():
    """The id numbering the CUDA kernel reproduces for the retention-time cut (cluster.py:418-429 combines two
    fcluster assignments through a non-injective map, so the ids themselves matter): the range-addition
    formulation equals scipy's fcluster on the linkage the reference builds, on random 1-D data with ties."""
    import scipy.cluster.hierarchy as sch

    rng = np.random.default_rng(11)
    for case in range(300):
        m = int(rng.integers(2, 60))
        mode = "ppm" if case % 3 == 0 else None
        base = rng.uniform(300, 1500) if mode == "ppm" else rng.uniform(0, 100)
        spread = base * 20e-6 if mode == "ppm" else 5.0
        v = base + rng.normal(0, spread * rng.uniform(0.2, 4.0), m)
        if case % 4 == 0:
            v[rng.integers(0, m, m // 3)] = v[0]  # ties
        if case % 5 == 0:
            v = np.round(v, 3 if mode is None else 5)  # more ties
        v = np.sort(v)
        tol = 20.0 if mode == "ppm" else 5.0
        want = sch.fcluster(odb.linkage_1d(v, mode), tol, "distance").astype(np.int64) - 1
        got = odb.fcluster_ids_by_range_additions(v, tol, mode)
        assert np.array_equal(got, want), (case, v.tolist())
