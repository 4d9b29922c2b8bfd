"""End-to-end parity (partitions identical to the oracle), the facades, and
size-independent properties at BASELINE sizes."""
import functools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from falcon_b200 import pipeline, synth  # noqa: E402
from falcon_b200.cluster import cluster, spectrum  # noqa: E402
from oracle import dbscan as odb  # noqa: E402
from oracle import ivf as oivf  # noqa: E402
from oracle import vectorize as ovec  # noqa: E402
from tests import helpers  # noqa: E402


def _cpu(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("n,lo,hi", [(50000, 700.0, 3500.0), (20000, 1000.0, 1020.0)])
def test_end_to_end_exhaustive_partition_identical(n, lo, hi):
    """BASELINE config 0 size (50k): identical partitions in exhaustive mode."""
    sp = helpers.dataset(n, 42, lo, hi)
    labels, nc, _ = pipeline.cluster_host(sp, pipeline.Settings(exhaustive=True))
    o = helpers.oracle_pipeline(sp, exhaustive=True)
    ref = np.empty(n, np.int64)
    ref[o["order"]] = o["labels"]
    assert nc == ref.max() + 1
    assert odb.same_partition(labels, ref)


@pytest.mark.parametrize("low_dim,eps,tol,mode,mz_interval", [
    (200, 0.05, 20.0, "ppm", 1), (800, 0.30, 20.0, "ppm", 1), (400, 0.20, 0.02, "Da", 2), (100, 0.10, 50.0, "ppm", 1)])
def test_end_to_end_settings_sweep(low_dim, eps, tol, mode, mz_interval):
    """BASELINE config 4's axes (low_dim 200 / 400 / 800, eps 0.05 - 0.30) plus the Da tolerance and a
    coarser bucket interval, at a size the oracle checks exhaustively: identical partitions."""
    n = 12000
    sp = helpers.dataset(n, 44, 1000.0, 1012.0)
    s = pipeline.Settings(exhaustive=True, low_dim=low_dim, eps=eps, precursor_tol_mass=tol, precursor_tol_mode=mode,
                          mz_interval=mz_interval)
    labels, nc, _ = pipeline.cluster_host(sp, s)
    o = helpers.oracle_pipeline(sp, exhaustive=True, low_dim=low_dim, eps=eps, tol=tol, mode=mode,
                                mz_interval=mz_interval)
    ref = np.empty(n, np.int64)
    ref[o["order"]] = o["labels"]
    assert nc == ref.max() + 1 and nc > 100
    assert odb.same_partition(labels, ref)


@pytest.mark.parametrize("rt_tol", [12.0, 60.0])
def test_end_to_end_with_rt_tolerance(rt_tol):
    """``rt_tol`` set: retention-time filter of the neighbours (A.2) and the retention-time cut of the
    DBSCAN clusters (cluster.py:418-429); identical partitions, and different from the run without it."""
    n = 20000
    sp = helpers.dataset(n, 45, 1000.0, 1020.0)
    labels, nc, _ = pipeline.cluster_host(sp, pipeline.Settings(exhaustive=True, rt_tol=rt_tol))
    o = helpers.oracle_pipeline(sp, exhaustive=True, rt_tol=rt_tol)
    ref = np.empty(n, np.int64)
    ref[o["order"]] = o["labels"]
    assert nc == ref.max() + 1 and nc > 100
    assert odb.same_partition(labels, ref)
    plain, _, _ = pipeline.cluster_host(sp, pipeline.Settings(exhaustive=True))
    if rt_tol < 20:
        assert not odb.same_partition(labels, plain)


@pytest.mark.parametrize("exhaustive", [False, True])
def test_sync_free_steady_state_equals_the_stage_by_stage_run(exhaustive):
    """Batches of one size through ONE HotPath: from the second batch on the stages are sized by the previous
    batch's counts and nothing is read back until the end of the step.  Labels must equal a fresh
    HotPath's (read-back after every stage), also when the new batch breaks the bounds learnt from the
    previous one (more buckets, more centroids, larger buckets, more pairs) and the step is redone."""
    n = 30000
    ranges = [(1000.0, 1010.0), (1000.0, 1010.0), (1002.0, 1012.5), (700.0, 3500.0), (1000.0, 1003.0),
              (1000.0, 1003.0), (900.0, 1100.0)]
    hp = pipeline.HotPath(pipeline.Settings(exhaustive=exhaustive))
    modes = []
    for i, (lo, hi) in enumerate(ranges):
        sp = synth.generate(n, 70 + i, mass_range=(lo, hi))
        d = helpers.to_device(sp, hp.device)
        fresh = pipeline.HotPath(pipeline.Settings(exhaustive=exhaustive, speculate=False))
        ref, nc_ref = fresh.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
        before = getattr(hp, "spec_misses", 0)
        speculated = hp._spec_for(n, False) is not None
        if i % 2:
            labels, nc = hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
        else:  # the host API shares the learnt bounds
            out = torch.empty(n, dtype=torch.int32).pin_memory()
            labels, nc = hp.run_host(*(t.cpu().pin_memory() for t in (d["mz"], d["intensity"], d["indptr"],
                                                                        d["precursor_mz"], d["charge"])), labels_out=out)
            torch.cuda.synchronize()
            assert np.array_equal(out.numpy(), _cpu(ref))
        assert nc == nc_ref and torch.equal(labels, ref), (i, lo, hi)
        modes.append("first" if not speculated else ("redo" if getattr(hp, "spec_misses", 0) > before else "ok"))
    assert modes[0] == "first" and modes[1] == "ok" and "redo" in modes[2:] and modes.count("ok") >= 3, modes


def test_end_to_end_default_nprobe_own_centroids_against_the_oracles():
    """Default n_probe with NO shared state: the device trains its own index, the oracle trains its own
    (`oracle/ivf.py:kmeans_train`; same conventions, but float32 scores summed in another order, so near-ties can
    fall differently and the indexes need not be identical).  What north_star asks of the approximate mode holds
    between the two: >= 0.99 of the oracle's within-eps neighbours are found (and vice versa), and the
    partitions agree almost everywhere (adjusted Rand index)."""
    from sklearn.metrics import adjusted_rand_score

    n = 30000
    sp = helpers.dataset(n, 48, 1000.0, 1010.0)  # buckets of ~1 500 rows: 32 lists, 4 probes
    h = pipeline.HotPath(pipeline.Settings(exhaustive=False))
    d = helpers.to_device(sp, h.device)
    labels, nc, keep = h.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    assert int(_cpu(keep["ivf"].nlist)[: keep["buckets"].n_buckets].max()) >= 16
    o = helpers.oracle_pipeline(sp, exhaustive=False)
    ref = o["csr_cut"]
    g = keep["graph"]
    gi, gp = _cpu(g.indices), _cpu(g.indptr)
    hit = sum(len(set(gi[gp[q]: gp[q + 1]]) & set(ref.indices[ref.indptr[q]: ref.indptr[q + 1]])) for q in range(n))
    assert hit / ref.nnz >= 0.99 and hit / g.nnz >= 0.99
    got = _cpu(keep["sorted_labels"]).astype(np.int64)
    want = np.asarray(o["labels"], np.int64)
    # noise as singletons for the index: -1 is not a cluster
    a = np.where(got >= 0, got, got.max() + 1 + np.arange(n))
    b = np.where(want >= 0, want, want.max() + 1 + np.arange(n))
    assert adjusted_rand_score(a, b) >= 0.99
    assert abs(nc - (int(want.max()) + 1)) <= 0.01 * nc


def test_end_to_end_default_nprobe_shared_centroids():
    sp = helpers.dataset(20000, 43, 1000.0, 1020.0)
    h = pipeline.HotPath(pipeline.Settings(exhaustive=False))
    d = helpers.to_device(sp, h.device)
    labels, nc, keep = h.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    ivf, b = keep["ivf"], keep["buckets"]
    nlist, cptr, cents = _cpu(ivf.nlist), _cpu(ivf.centroid_ptr), _cpu(ivf.centroids)
    assert (nlist > 0).sum() > 10
    shared = [cents[cptr[i]: cptr[i] + nlist[i]] if nlist[i] else None for i in range(b.n_buckets)]
    o = helpers.oracle_pipeline(sp, exhaustive=False, centroids=shared, vectors=_cpu(keep["x"]))
    ref = np.empty(len(sp), np.int64)
    ref[o["order"]] = o["labels"]
    assert odb.same_partition(_cpu(labels), ref)


def test_facade_to_vector_parallel_and_errors():
    sp = helpers.dataset(2000, 3)
    vec_len, lo, hi = spectrum.get_dim(101.0, 1500.0, 0.05)
    table = spectrum.hash_lookup(vec_len, 400)
    v = spectrum.to_vector_parallel(sp.as_dicts(), dim=400, min_mz=lo, max_mz=hi, bin_size=0.05,
                                    hash_lookup=table, norm=True)
    np.testing.assert_allclose(v, helpers.oracle_vectors(sp), rtol=0, atol=1e-6)
    with pytest.raises(NotImplementedError):
        spectrum.to_vector_parallel(sp.as_dicts()[:4], 400, lo, hi, 0.05, hash_lookup=np.zeros(vec_len, np.uint32))
    assert spectrum.to_vector_parallel([], 400, lo, hi, 0.05).shape == (0, 400)


def test_facade_compute_pairwise_distances_and_generate_clusters(tmp_path):
    import joblib

    sp = helpers.dataset(6000, 5, 1000.0, 1010.0)
    order, bptr, _ = oivf.bucket_sort(sp.precursor_mz, sp.precursor_charge)
    dicts = sp.as_dicts()
    files = []
    for i, (s, e) in enumerate(zip(bptr[:-1], bptr[1:])):  # per-bucket files like the reference's work_dir
        f = tmp_path / f"bucket_{i}.pkl"
        joblib.dump([dicts[j] for j in order[s:e]], f)
        files.append(str(f))
    vec_len, lo, hi = spectrum.get_dim(101.0, 1500.0, 0.05)
    vectorize = functools.partial(spectrum.to_vector_parallel, dim=400, min_mz=lo, max_mz=hi, bin_size=0.05,
                                  hash_lookup=None, norm=True)
    mat, meta = cluster.compute_pairwise_distances(len(sp), files, None, vectorize, 20.0, "ppm", None, 64, 128,
                                                   2 ** 16, 32, eps=0.1, exhaustive=True)
    o = helpers.oracle_pipeline(sp, exhaustive=True)
    assert np.array_equal(mat.indptr, o["csr_cut"].indptr) and np.array_equal(mat.indices, o["csr_cut"].indices)
    assert mat.dtype == np.float32 and mat.shape == (len(sp), len(sp))
    assert np.array_equal(meta["precursor_mz"].values, o["sorted"].precursor_mz)
    labels = cluster.generate_clusters(mat, 0.1, meta["precursor_mz"].values, None, 20.0, "ppm", None)
    assert odb.same_partition(labels, o["labels"])
    # eps=None: the full n_neighbors matrix like the reference's (built bucket range by bucket range), with
    # get_cluster_representatives on it
    full, _ = cluster.compute_pairwise_distances(len(sp), files, None, vectorize, 20.0, "ppm", None, 64, 128,
                                                 2 ** 16, 32, eps=None, exhaustive=True)
    assert np.array_equal(full.indptr, o["csr"].indptr) and np.array_equal(full.indices, o["csr"].indices)
    np.testing.assert_allclose(full.data, o["csr"].data, rtol=0, atol=1e-5)
    assert full.nnz > 2 * mat.nnz  # beyond-eps neighbours inside the precursor tolerance are there as well
    from oracle import medoids as omed

    reps = cluster.get_cluster_representatives(labels, full.indptr, full.indices, full.data)
    assert np.array_equal(reps, omed.cluster_medoids(o["csr"].data, o["csr"].indices, o["csr"].indptr, labels))
    # the oracle's own full matrix through the GPU generate_clusters
    labels2 = cluster.generate_clusters(o["csr"], 0.1, o["sorted"].precursor_mz, None, 20.0, "ppm")
    assert odb.same_partition(labels2, o["labels"])
    with pytest.raises(ValueError, match="tolerance mode"):
        cluster.generate_clusters(mat, 0.1, meta["precursor_mz"].values, None, 20.0, "mDa")
    with pytest.raises(ValueError):
        cluster.compute_pairwise_distances(len(sp) + 1, files, None, vectorize, 20.0, "ppm", None, 64, 128, 1, 32)


def test_empty_and_tiny_inputs():
    h = pipeline.HotPath(pipeline.Settings(exhaustive=True))
    e = synth.generate(0)
    labels, nc, _ = pipeline.cluster_host(e)
    assert labels.shape == (0,) and nc == 0
    one = synth.generate(1, 5)
    labels, nc, _ = pipeline.cluster_host(one, pipeline.Settings(exhaustive=True))
    assert labels.tolist() == [-1] and nc == 0
    two = one.take(np.array([0, 0]))
    labels, nc, _ = pipeline.cluster_host(two, pipeline.Settings(exhaustive=True))
    assert labels.tolist() == [0, 0] and nc == 1


def test_full_size_properties_1m():
    """BASELINE config 1 (1M spectra, defaults): properties that need no oracle."""
    n = 1_000_000
    sp = helpers.dataset(n, 42)
    s = pipeline.Settings()
    labels, nc, _ = pipeline.cluster_host(sp, s)
    labels2, nc2, _ = pipeline.cluster_host(sp, s)
    assert nc == nc2 and np.array_equal(labels, labels2)  # deterministic
    m = labels >= 0
    assert nc > n // 20 and labels.max() == nc - 1
    sizes = np.bincount(labels[m], minlength=nc)
    assert sizes.min() >= 2  # min_samples
    # precursor tolerance: complete linkage span of every cluster within 20 ppm of its lightest member
    lo = np.full(nc, np.inf)
    hi = np.zeros(nc)
    np.minimum.at(lo, labels[m], sp.precursor_mz[m])
    np.maximum.at(hi, labels[m], sp.precursor_mz[m])
    assert ((hi - lo) / lo * 1e6 <= 20.0).all()
    # one charge per cluster (buckets never mix charges)
    zmin = np.full(nc, 99)
    zmax = np.zeros(nc, np.int64)
    np.minimum.at(zmin, labels[m], sp.precursor_charge[m])
    np.maximum.at(zmax, labels[m], sp.precursor_charge[m])
    assert (zmin == zmax).all()
    # clusters are pure w.r.t. the generating templates on this synthetic set
    tmin = np.full(nc, np.iinfo(np.int64).max)
    tmax = np.zeros(nc, np.int64)
    np.minimum.at(tmin, labels[m], sp.template[m])
    np.maximum.at(tmax, labels[m], sp.template[m])
    assert (tmin == tmax).mean() > 0.999
    # exhaustive mode finds a superset of the default-n_probe neighbours: never fewer clustered spectra
    labels_ex, _, _ = pipeline.cluster_host(sp, pipeline.Settings(exhaustive=True))
    assert (labels_ex >= 0).sum() >= m.sum()


def test_full_size_exhaustive_parity_1m():
    """BASELINE config 2 (1M spectra, n_probe = nlist): neighbour lists, distances and the
    partition against the oracle at full size (the oracle spreads its buckets over the host cores)."""
    import os

    from oracle import pipeline as opipe

    n = 1_000_000
    sp = helpers.dataset(n, 42)
    h = pipeline.HotPath(pipeline.Settings(exhaustive=True))
    d = helpers.to_device(sp, h.device)
    labels, nc, keep = h.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    ref_labels, _, ref = opipe.run(sp, exhaustive=True, n_jobs=len(os.sched_getaffinity(0)), f32_gemm=False)
    g = keep["graph"]
    cut = oivf.eps_cut(ref, 0.1)
    assert np.array_equal(_cpu(g.indptr), cut.indptr)
    assert np.array_equal(_cpu(g.indices), cut.indices)  # same neighbours in the same order
    np.testing.assert_allclose(_cpu(g.dist), cut.data, rtol=0, atol=1e-5)
    assert odb.same_partition(_cpu(labels), ref_labels)
    assert nc == int(ref_labels.max()) + 1


@pytest.mark.gpu
def test_cli_mgf_to_csv_matches_oracle(tmp_path):
    """BASELINE configs[0] in miniature, through the `falcon` command: MGF in -> preprocessing ->
    per-charge clustering -> CSV (+ representatives MGF).  The partition must equal the oracle's
    (oracle preprocessing + oracle pipeline, exhaustive mode) on the same file."""
    import pandas as pd

    from falcon_b200 import falcon as fmain
    from falcon_b200.ms_io import mgf_io
    from oracle import preprocess as opre

    sp = synth.generate(4000, 17, mass_range=(1000.0, 1015.0))
    dicts = sp.as_dicts()
    for d in dicts[::7]:  # a dominant peak between get_dim's lower bound (100.95) and the raw setting (101.0)
        d["mz"] = np.r_[np.float32(100.97), d["mz"]].astype(np.float32)
        d["intensity"] = np.r_[np.float32(50.0 * d["intensity"].max()), d["intensity"]].astype(np.float32)
    path, out = str(tmp_path / "in.mgf"), str(tmp_path / "res")
    mgf_io.write_spectra(path, dicts)
    assert fmain.main([path, out, "--exhaustive", "--export_representatives"]) == 0
    with open(out + ".csv") as fh:
        header = [l for l in fh if l.startswith("#")]
    assert header[0].startswith("# falcon version") and any(l.startswith("# eps = 0.100") for l in header)
    df = pd.read_csv(out + ".csv", comment="#")
    assert list(df.columns) == ["filename", "spectrum_id", "precursor_charge", "precursor_mz", "retention_time", "cluster"]
    # rows in natural order of (filename, spectrum_id) like natsort (falcon.py:206-208): synth:2 before synth:10
    nums = [int(t.split(":")[1]) for t in df["spectrum_id"]]
    assert nums == sorted(nums) and nums != sorted(nums, key=str)
    # a second run without --overwrite refuses (return code 1) and leaves the files alone
    assert fmain.main([path, out, "--exhaustive"]) == 1
    # oracle: same preprocessing, same path
    raw, ids, _ = mgf_io.read_mgf(path)
    # falcon.py:120-133: the m/z window the spectra are restricted to is get_dim's, not the raw setting
    _, lo, hi = ovec.get_dim(101.0, 1500.0, 0.05)
    assert (lo, hi) != (101.0, 1500.0)
    valid, mz, it, indptr = opre.process_spectra(raw, min_peaks=5, min_mz_range=250.0, mz_min=lo, mz_max=hi,
                                                 remove_precursor_tolerance=1.5, min_intensity=0.01, max_peaks_used=50,
                                                 scaling=None)
    keep = np.flatnonzero(valid)
    assert sorted(df["spectrum_id"]) == sorted(ids[i] for i in keep)
    proc = synth.SpectrumSet(mz, it, indptr, raw.precursor_mz, raw.precursor_charge, raw.retention_time).take(keep)
    o = helpers.oracle_pipeline(proc, exhaustive=True)
    ref = np.empty(len(proc), np.int64)
    ref[o["order"]] = o["labels"]
    by_id = dict(zip(df["spectrum_id"], df["cluster"]))
    got = np.array([by_id[ids[i]] for i in keep])
    assert odb.same_partition(got, ref)
    # labels of the two charges are disjoint and consecutive
    lab = df[df["cluster"] >= 0]
    assert lab.groupby("cluster")["precursor_charge"].nunique().max() == 1
    assert sorted(lab["cluster"].unique()) == list(range(lab["cluster"].nunique()))
    # one representative per cluster, each a member of its cluster
    reps = list(mgf_io.get_spectra(out + ".mgf"))
    assert len(reps) == lab["cluster"].nunique()
    # --overwrite runs again; every unclustered spectrum can get its own id (development-head convention)
    assert fmain.main([path, out, "--exhaustive", "--overwrite", "--singletons_as_clusters"]) == 0
    df2 = pd.read_csv(out + ".csv", comment="#")
    assert (df2["cluster"] >= 0).all() and df2["cluster"].nunique() == lab["cluster"].nunique() + int((df["cluster"] < 0).sum())
    assert (df2["cluster"][df["cluster"] >= 0] == df["cluster"][df["cluster"] >= 0]).all()


@pytest.mark.gpu
def test_facade_process_spectrum():
    from oracle import preprocess as opre

    mz = np.float32([100, 150, 200.5, 300, 400, 500, 600, 700])
    it = np.float32([1, 5, 2, 9, 9, 0.05, 3, 4])
    raw = {"identifier": "x", "precursor_mz": 300.0, "precursor_charge": 2, "mz": mz, "intensity": it,
           "retention_time": 1.0, "filename": "f.mgf"}
    kw = dict(min_peaks=3, min_mz_range=100.0, mz_min=101, mz_max=1500, remove_precursor_tolerance=1.5,
              min_intensity=0.01, max_peaks_used=4, scaling="rank")
    got = spectrum.process_spectrum(raw, **kw)
    ref_mz, ref_it = opre.process_spectrum(mz, it, 300.0, 2, **kw)
    assert got["identifier"] == "x" and got["filename"] == "f.mgf" and got["precursor_charge"] == 2
    assert np.array_equal(got["mz"], ref_mz) and np.allclose(got["intensity"], ref_it, atol=1e-7)
    assert spectrum.process_spectrum(raw, **{**kw, "min_peaks": 5}) is None
    with pytest.raises(ValueError):
        spectrum.process_spectrum(raw, **{**kw, "scaling": "cube"})
