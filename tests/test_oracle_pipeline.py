"""Self-consistency of the oracle's unpinned stages (buckets / IVF / DBSCAN)."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import dbscan as odb
from oracle import ivf as oivf
from tests import helpers


def test_nlist_rule():
    assert oivf.n_list_rule(99) == 0
    assert oivf.n_list_rule(100) == 2
    assert oivf.n_list_rule(155) == 2 and oivf.n_list_rule(156) == 4
    assert oivf.n_list_rule(39 * 64) == 64 and oivf.n_list_rule(39 * 64 - 1) == 32
    assert oivf.n_list_rule(10 ** 6) == 2 ** 16
    assert oivf.n_probe_rule(64, 32) == 8 and oivf.n_probe_rule(1024, 32) == 32
    assert oivf.n_probe_rule(2, 32) == 1 and oivf.n_probe_rule(64, 32, exhaustive=True) == 64


def test_bucket_rule_and_sort():
    mz = np.array([500.2, 500.7, 500.2004, 334.1, 500.1])
    z = np.array([2, 2, 2, 3, 2])
    b = oivf.bucket_ids(mz, z)
    assert b.tolist() == [int(round((m - 1.00794) * c / 1.0005079)) for m, c in zip(mz, z)]
    order, ptr, keys = oivf.bucket_sort(mz, z)
    ks = ((z.astype(np.int64) << 24) | b)[order]
    assert (np.diff(ks) >= 0).all() and ptr[0] == 0 and ptr[-1] == 5
    for s, e in zip(ptr[:-1], ptr[1:]):
        assert (np.diff(mz[order][s:e]) >= 0).all()


def test_exhaustive_equals_brute_force():
    sp = helpers.dataset(3000, 3, 1000.0, 1010.0)
    o = helpers.oracle_pipeline(sp, exhaustive=True)
    x, mzs, bptr = o["x"], o["sorted"].precursor_mz, o["bucket_ptr"]
    sim = (x.astype(np.float64) @ x.astype(np.float64).T).astype(np.float32)
    csr = o["csr"]
    assert np.diff(bptr).max() >= 100
    for q in range(0, 3000, 97):
        b = np.searchsorted(bptr, q, "right") - 1
        s, e = bptr[b], bptr[b + 1]
        cand = np.arange(s, e)
        order = cand[np.argsort(-sim[q, s:e], kind="stable")][:128]
        ok = np.abs(mzs[q] - mzs[order]) / mzs[order] * 1e6 < helpers.TOL
        exp = order[ok][:64]
        got = csr.indices[csr.indptr[q]: csr.indptr[q + 1]]
        assert np.array_equal(got, exp)
        np.testing.assert_array_equal(csr.data[csr.indptr[q]: csr.indptr[q + 1]],
                                      np.maximum(np.float32(1) - sim[q, exp], 0))
        assert q in got


def test_ivf_is_subset_and_recall_reasonable():
    sp = helpers.dataset(3000, 3, 1000.0, 1010.0)
    ex = helpers.oracle_pipeline(sp, exhaustive=True)
    iv = helpers.oracle_pipeline(sp, exhaustive=False)
    a, b = ex["csr_cut"], iv["csr_cut"]
    hit = tot = 0
    for q in range(3000):
        ea = set(a.indices[a.indptr[q]: a.indptr[q + 1]].tolist())
        eb = set(b.indices[b.indptr[q]: b.indptr[q + 1]].tolist())
        assert eb <= ea
        hit += len(eb)
        tot += len(ea)
    assert hit / tot > 0.5


@settings(max_examples=150, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_dbscan_min_ancestor_equals_sklearn(n, seed):
    """dbscan_inner == rank of the minimum-index core ancestor."""
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, 5, n)
    indptr = np.r_[0, np.cumsum(deg)]
    indices = np.concatenate([np.r_[i, rng.integers(0, n, d - 1)] for i, d in enumerate(deg)])
    data = rng.random(indices.shape[0]).astype(np.float32) * 0.2
    data[indptr[:-1]] = 0
    a = odb.dbscan_sklearn(data, indices, indptr, 0.1)
    b = odb.dbscan_min_ancestor(data, indices, indptr, 0.1)
    assert np.array_equal(a, b)


def test_oracle_clusters_are_pure_on_synthetic_data():
    sp = helpers.dataset(4000, 11)
    o = helpers.oracle_pipeline(sp, exhaustive=True)
    lab, tmpl = o["labels"], o["sorted"].template
    assert lab.max() > 300
    for l in np.unique(lab[lab >= 0])[:200]:
        assert np.unique(tmpl[lab == l]).shape[0] == 1


def test_medoid_oracle_hand_example():
    """SURVEY A.5 on a 5-member cluster + a pair + noise: eligibility needs more than
    a quarter of the cluster in the row; first minimum wins; pairs give the first member."""
    import scipy.sparse as ss

    from oracle import medoids as omed

    labels = np.array([0, 0, 0, 0, 0, 1, 1, -1], np.int32)
    rows = {
        0: {0: 0.0, 1: 0.4, 2: 0.4},          # mean 0.2667 over 3 members
        1: {1: 0.0, 0: 0.1, 2: 0.1, 3: 0.1},  # mean 0.075 -> the medoid
        2: {2: 0.0, 7: 0.01},                 # only itself present: 1 <= 5 / 4 -> not eligible
        3: {3: 0.0, 1: 0.15, 5: 0.0},         # other cluster's member ignored: mean 0.075, ties lose to row 1
        4: {4: 0.0, 0: 0.3},                  # mean 0.15
        5: {5: 0.0, 6: 0.05},
        6: {6: 0.0, 5: 0.05},
        7: {7: 0.0},
    }
    m = ss.lil_matrix((8, 8), dtype=np.float32)
    data, indices, indptr = [], [], [0]
    for r in range(8):
        for c, v in rows[r].items():
            indices.append(c)
            data.append(v)
        indptr.append(len(indices))
    got = omed.cluster_medoids(np.float32(data), np.int32(indices), np.int64(indptr), labels)
    assert got.tolist() == [1, 5]
    assert omed.cluster_medoids(np.zeros(0, np.float32), np.zeros(0, np.int32), np.zeros(1, np.int64),
                                np.zeros(0, np.int32)).shape == (0,)


def test_preprocess_oracle_hand_example():
    """process_spectrum on a worked example: window, precursor peaks of charge 2 and 1,
    1 % base-peak threshold, top 4, rank scaling, unit norm."""
    from oracle import preprocess as opre

    mz = np.float32([100, 150, 200.5, 300, 400, 500, 600, 700])
    it = np.float32([1, 5, 2, 9, 9, 0.05, 3, 4])
    kw = dict(min_peaks=3, min_mz_range=100.0, mz_min=101, mz_max=1500, remove_precursor_tolerance=1.5,
              min_intensity=0.01, max_peaks_used=4)
    m, i = opre.process_spectrum(mz, it, 300.0, 2, scaling="rank", **kw)
    assert m.tolist() == [150.0, 200.5, 400.0, 700.0]
    np.testing.assert_allclose(i, np.float32([3, 1, 4, 2]) / np.sqrt(np.float32(30)), atol=1e-7)
    assert opre.process_spectrum(mz, it, 300.0, 2, scaling=None, **{**kw, "min_peaks": 5}) is None
    m, i = opre.process_spectrum(mz, it, 300.0, 2, scaling="root", **{**kw, "max_peaks_used": None})
    assert m.tolist() == [150.0, 200.5, 400.0, 700.0] and abs(float((i.astype(np.float64) ** 2).sum()) - 1) < 1e-6


def test_oracle_kmeans_against_sklearn_kmeans():
    """An independent second source for the IVF trainer the oracle restates (faiss itself is not installable
    here): scikit-learn's Lloyd k-means, started from the same rows, on the same unit vectors.  For unit
    vectors the Euclidean objective is 2 - 2 * (mean inner product with the own centroid direction), so the
    two must reach the same quality; the assignments of the oracle's spherical variant and sklearn's centroids
    (normalised) must largely agree."""
    from sklearn.cluster import KMeans

    sp = helpers.dataset(4000, 9, 1000.0, 1001.0)
    x = helpers.oracle_vectors(sp)
    order, bptr, _ = oivf.bucket_sort(sp.precursor_mz, sp.precursor_charge)
    sizes = np.diff(bptr)
    b = int(np.argmax(sizes))
    xb = x[order[bptr[b]: bptr[b + 1]]]
    k = oivf.n_list_rule(xb.shape[0])
    assert k >= 16
    cent = oivf.kmeans_train(xb, k, 10)
    init = xb[(np.arange(k, dtype=np.int64) * xb.shape[0]) // k]
    km = KMeans(n_clusters=k, init=init, n_init=1, max_iter=10, algorithm="lloyd", tol=0.0).fit(xb.astype(np.float64))
    sk = km.cluster_centers_ / np.linalg.norm(km.cluster_centers_, axis=1, keepdims=True)

    def quality(c):
        return float((xb.astype(np.float64) @ c.T).max(axis=1).mean())

    q_oracle, q_sklearn, q_init = quality(cent.astype(np.float64)), quality(sk), quality(init.astype(np.float64))
    assert q_oracle > q_init + 0.02 and q_sklearn > q_init + 0.02  # training does something on this data
    assert abs(q_oracle - q_sklearn) < 0.03 * q_sklearn  # spherical vs Euclidean means: not the same fixed point
    # same rows grouped together by both solutions (adjusted Rand index of the two assignments)
    from sklearn.metrics import adjusted_rand_score

    a = oivf.assign_lists(xb, cent)
    s = np.argmax(xb.astype(np.float64) @ sk.T, axis=1)
    assert adjusted_rand_score(a, s) > 0.5  # 64 lists over ~600 templates: many equally good groupings
