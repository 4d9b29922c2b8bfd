"""Stage-by-stage parity of the CUDA path (through the C ABI) against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from falcon_b200 import pipeline, synth  # noqa: E402
from oracle import dbscan as odb  # noqa: E402
from oracle import hashing  # noqa: E402
from oracle import ivf as oivf  # noqa: E402
from oracle import vectorize as ovec  # noqa: E402
from tests import helpers  # noqa: E402


@pytest.fixture(scope="module")
def hp():
    return pipeline.HotPath(pipeline.Settings(exhaustive=True))


def _cpu(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------ stage 1
def test_hash_table_bit_exact(hp):
    got = _cpu(hp.hash_table()).view(np.uint32)
    assert np.array_equal(got, hashing.hash_lookup(hp.vec_len, 400))


def test_vectorize_golden_bins_bit_exact(hp, golden_dir):
    g = np.load(os.path.join(golden_dir, "binning.npz"))
    dev = hp.device
    mz, inten, indptr = (torch.from_numpy(g[k]).to(dev) for k in ("mz", "intensity", "indptr"))
    x, xb, hidx = hp.vectorize(mz, inten, indptr, want_hash_idx=True)
    table = hashing.hash_lookup(hp.vec_len, 400)
    assert np.array_equal(_cpu(hidx), table[g["bins"]].astype(np.int32))  # bins from the reference itself
    ref = ovec.to_vector(g["mz"], g["intensity"], g["indptr"], hp.min_mz, 0.05, hp.vec_len, 400)
    np.testing.assert_allclose(_cpu(x), ref, rtol=0, atol=1e-6)
    assert (_cpu(x) != ref).mean() < 1e-5
    assert np.array_equal(_cpu(xb.view(torch.int16)).view(np.uint16)[:, :400], ovec.to_bf16_bits(_cpu(x)))


@pytest.mark.parametrize("low_dim", [400, 200, 800, 100])
def test_vectorize_synthetic(low_dim):
    h = pipeline.HotPath(pipeline.Settings(low_dim=low_dim))
    sp = helpers.dataset(5000, 1)
    d = helpers.to_device(sp, h.device)
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], want_hash_idx=True)
    x, xb, hidx = v
    ref, ref_idx = helpers.oracle_vectors(sp, low_dim, return_hash_idx=True)
    assert np.array_equal(_cpu(hidx), ref_idx)
    # sparse (ELL) copy: distinct non-zero columns + values, zero padded, population in ell_nnz
    ei, ev = _cpu(v.ell_idx).view(np.uint16), _cpu(v.ell_val)
    assert v.ell_width % 8 == 0 and v.ell_width >= np.diff(sp.indptr).max()
    dense = np.zeros_like(_cpu(x))
    rows = np.repeat(np.arange(len(sp)), v.ell_width).reshape(ei.shape)
    np.add.at(dense, (rows, ei), ev)
    assert np.array_equal(dense, _cpu(x))
    nnz = (ev != 0).sum(axis=1)
    assert np.array_equal(nnz, (_cpu(x) != 0).sum(axis=1))
    assert np.array_equal(nnz, _cpu(v.ell_nnz).view(np.uint16))
    for r in range(0, len(sp), 500):
        assert len(set(ei[r, : nnz[r]].tolist())) == nnz[r] and (ev[r, nnz[r]:] == 0).all()
    np.testing.assert_allclose(_cpu(x), ref, rtol=0, atol=1e-6)
    xb_bits = _cpu(xb.view(torch.int16)).view(np.uint16)
    assert np.array_equal(xb_bits[:, :low_dim], ovec.to_bf16_bits(_cpu(x)))
    assert (xb_bits[:, low_dim:] == 0).all()
    # un-normalised and permuted
    order = torch.randperm(len(sp), device=h.device).to(torch.int32)
    x2, _, _ = h.vectorize(d["mz"], d["intensity"], d["indptr"], order=order, want_bf16=False, norm=False)
    ref2 = helpers.oracle_vectors(sp, low_dim, norm=False)[_cpu(order)]
    assert np.array_equal(_cpu(x2), ref2)


def test_vectorize_edge_cases(hp):
    dev = hp.device
    # spectrum 0 empty, spectrum 1 has out-of-range peaks and 3 peaks in one bin, spectrum 2 has 70 peaks
    mz = np.r_[np.float32([50.0, 200.01, 200.02, 200.03, 1600.0]),
               np.sort(np.random.default_rng(0).uniform(101, 1500, 70)).astype(np.float32)]
    inten = np.random.default_rng(1).random(mz.shape[0]).astype(np.float32)
    indptr = np.array([0, 0, 5, 75], np.int64)
    x, _, hidx = hp.vectorize(*(torch.from_numpy(a).to(dev) for a in (mz, inten, indptr)), want_hash_idx=True)
    ref, ref_idx = ovec.to_vector(mz, inten, indptr, hp.min_mz, 0.05, hp.vec_len, 400, return_hash_idx=True)
    assert np.array_equal(_cpu(hidx), ref_idx) and ref_idx[0] == -1 and ref_idx[4] == -1
    np.testing.assert_allclose(_cpu(x), ref, rtol=0, atol=1e-6)
    assert (_cpu(x)[0] == 0).all()
    e = torch.empty(0, device=dev)
    x0, _, _ = hp.vectorize(e.float(), e.float(), torch.zeros(1, dtype=torch.int64, device=dev))
    assert x0.shape == (0, 400)


# ------------------------------------------------------------------ buckets
@pytest.mark.parametrize("mz_interval", [1, 4])
def test_bucket_sort(mz_interval):
    h = pipeline.HotPath(pipeline.Settings(mz_interval=mz_interval))
    sp = helpers.dataset(20000, 5)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"], d["rt"])
    order, bptr, keys = oivf.bucket_sort(sp.precursor_mz, sp.precursor_charge, mz_interval)
    assert b.n_buckets == bptr.shape[0] - 1
    assert np.array_equal(_cpu(b.bucket_ptr), bptr)
    assert np.array_equal(_cpu(b.order), order)
    assert np.array_equal(_cpu(b.mz), sp.precursor_mz[order])
    assert np.array_equal(_cpu(b.rt), sp.retention_time[order])
    assert np.array_equal(_cpu(b.key).view(np.uint32)[bptr[:-1]], keys)


# ------------------------------------------------------------------ scan
def _pairs(h, xb, buckets, n, impl, thr, cap=1 << 22):
    import ctypes as C

    from falcon_b200._lib import check, lib, ptr

    pairs = torch.empty(cap, dtype=torch.int64, device=h.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=h.device)
    ws = torch.empty(lib.flc_scan_workspace_bytes(n, buckets.n_buckets) + 256, dtype=torch.uint8, device=h.device)
    check(lib.flc_scan_pairs(ptr(xb), xb.stride(0), n, h.s.low_dim, ptr(buckets.bucket_ptr), buckets.n_buckets,
                             None, None, 0, None, thr, impl, ptr(pairs), cap, ptr(cnt), ptr(ws), ws.numel(),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    k = int(cnt.item())
    assert k <= cap
    p = _cpu(pairs[:k]).view(np.uint64)
    return set(zip((p >> np.uint64(32)).astype(np.int64).tolist(), (p & np.uint64(0xFFFFFFFF)).astype(np.int64).tolist()))


@pytest.mark.parametrize("n,lo,hi,low_dim", [(4000, 700.0, 3500.0, 400), (6000, 1000.0, 1012.0, 400),
                                             (3000, 1000.0, 1003.0, 200), (3000, 1000.0, 1003.0, 800)])
def test_scan_pairs_tc_and_simt(n, lo, hi, low_dim):
    """Both scan kernels return every pair whose exact ip clears 1 - eps, nothing
    outside the bucket, nothing below the bf16 threshold by more than rounding."""
    h = pipeline.HotPath(pipeline.Settings(low_dim=low_dim, exhaustive=True))
    sp = helpers.dataset(n, 9, lo, hi)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    x, xb = v.x, v.xb
    thr = h.scan_threshold()
    xf = _cpu(xb.float())[:, :low_dim].astype(np.float64)
    xe = _cpu(x).astype(np.float64)
    bptr = _cpu(b.bucket_ptr)
    must, may = set(), set()
    for s, e in zip(bptr[:-1], bptr[1:]):
        sb = xf[s:e] @ xf[s:e].T
        se = (xe[s:e] @ xe[s:e].T).astype(np.float32)
        for q, c in zip(*np.nonzero(np.maximum(np.float32(1) - se, 0) <= np.float32(helpers.EPS))):
            must.add((q + s, c + s))
        for q, c in zip(*np.nonzero(sb >= thr - 1e-3)):
            may.add((q + s, c + s))
    assert len(must) > n
    for impl in (1, 0):
        got = _pairs(h, xb, b, n, impl, thr)
        assert must <= got, f"impl {impl} missed {len(must - got)} pairs"
        assert got <= may, f"impl {impl} produced {len(got - may)} spurious pairs"


# ------------------------------------------------------------------ k-NN CSR, exhaustive
@pytest.mark.parametrize("n,lo,hi,scan_impl", [(5000, 700.0, 3500.0, 0), (6000, 1000.0, 1012.0, 0),
                                               (6000, 1000.0, 1012.0, 1)])
def test_knn_csr_exhaustive_exact(n, lo, hi, scan_impl):
    h = pipeline.HotPath(pipeline.Settings(exhaustive=True, scan_impl=scan_impl))
    sp = helpers.dataset(n, 13, lo, hi)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    x, xb = v.x, v.xb
    g = h.knn_graph(v, b)
    o = helpers.oracle_pipeline(sp, exhaustive=True, vectors=_cpu(x))
    ref = o["csr_cut"]
    assert np.array_equal(_cpu(g.indptr), ref.indptr)
    assert np.array_equal(_cpu(g.indices), ref.indices)  # neighbour sets AND order identical
    np.testing.assert_allclose(_cpu(g.dist), ref.data, rtol=0, atol=1e-5)  # north_star tolerance
    assert (_cpu(g.dist) != ref.data).mean() < 1e-4  # in fact bit-identical up to double rounding


def test_knn_csr_without_eps_cut_matches_full_reference_matrix():
    h = pipeline.HotPath(pipeline.Settings(exhaustive=True, eps_cut=False))
    sp = helpers.dataset(3000, 17)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    x, xb = v.x, v.xb
    g = h.knn_graph(v, b)
    ref = helpers.oracle_pipeline(sp, exhaustive=True, vectors=_cpu(x))["csr"]
    assert np.array_equal(_cpu(g.indptr), ref.indptr)
    assert np.array_equal(_cpu(g.indices), ref.indices)
    np.testing.assert_allclose(_cpu(g.dist), ref.data, rtol=0, atol=1e-5)


def test_knn_csr_da_mode_and_small_k():
    s = pipeline.Settings(exhaustive=True, precursor_tol_mass=0.01, precursor_tol_mode="Da", n_neighbors=3,
                          n_neighbors_ann=5, eps=0.3)
    h = pipeline.HotPath(s)
    sp = helpers.dataset(4000, 19, 1000.0, 1004.0)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    x, xb = v.x, v.xb
    g = h.knn_graph(v, b)
    order, bptr, _ = oivf.bucket_sort(sp.precursor_mz, sp.precursor_charge)
    mat, _ = oivf.compute_pairwise_distances(_cpu(x), sp.precursor_mz[order], None, bptr, 0.01, "Da", None,
                                             3, 5, 32, True)
    ref = oivf.eps_cut(mat, 0.3)
    assert np.array_equal(_cpu(g.indptr), ref.indptr)
    assert np.array_equal(_cpu(g.indices), ref.indices)


# ------------------------------------------------------------------ IVF
def _drop_ell(v):
    import dataclasses

    return dataclasses.replace(v, ell_idx=None, ell_val=None, ell_nnz=None, ell_width=0)


@pytest.mark.parametrize("sparse", [True, False])
def test_ivf_shared_centroids_exact_and_recall(sparse):
    """Default n_probe with centroids shared between the CUDA path and the oracle:
    neighbour sets identical (north_star asks recall >= 0.99).  sparse=True uses the
    ELL rows (final assignment fused into the trainer), sparse=False re-assigns the
    dense rows to the trained centroids with flc_ivf_assign."""
    h = pipeline.HotPath(pipeline.Settings(exhaustive=False))
    sp = helpers.dataset(8000, 23, 1000.0, 1008.0)
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    x, xb = v.x, v.xb
    ivf = h.build_ivf(v, b)
    if not sparse:
        with pytest.raises(ValueError):
            h.build_ivf(_drop_ell(v), b)
        v = _drop_ell(v)
        ivf = h.build_ivf(v, b, centroids=ivf.centroids)
    assert ivf.total_centroids > 0
    nlist, cptr = _cpu(ivf.nlist), _cpu(ivf.centroid_ptr)
    bptr = _cpu(b.bucket_ptr)
    assert nlist[:-1].tolist() == [oivf.n_list_rule(int(e - s)) for s, e in zip(bptr[:-1], bptr[1:])]
    cents = _cpu(ivf.centroids)
    shared = [cents[cptr[i]: cptr[i] + nlist[i]] if nlist[i] else None for i in range(b.n_buckets)]
    # coarse assignment / probes equal the oracle's
    xs = _cpu(x)
    lid, probes = _cpu(ivf.list_id), _cpu(ivf.probes)
    for i in range(b.n_buckets):
        if nlist[i]:
            s, e = bptr[i], bptr[i + 1]
            assert np.array_equal(lid[s:e], oivf.assign_lists(xs[s:e], shared[i]))
            p = oivf.n_probe_rule(int(nlist[i]), 32)
            assert np.array_equal(probes[s:e, :p], oivf.probe_lists(xs[s:e], shared[i], p))
            assert (probes[s:e, p:] == -1).all()
    g = h.knn_graph(v, b, ivf)
    o = helpers.oracle_pipeline(sp, exhaustive=False, centroids=shared, vectors=xs)
    ref = o["csr_cut"]
    gi, gp = _cpu(g.indices), _cpu(g.indptr)
    hit = sum(len(set(gi[gp[q]: gp[q + 1]]) & set(ref.indices[ref.indptr[q]: ref.indptr[q + 1]]))
              for q in range(len(sp)))
    assert hit / max(ref.nnz, 1) >= 0.99
    assert np.array_equal(gp, ref.indptr) and np.array_equal(gi, ref.indices)


def test_ivf_more_than_32_probes():
    """n_probe = 64 on a bucket of > 20 000 rows (512 lists, min(ceil(512 / 8), 64) = 64 probes): probe lists
    from the trainer's final assignment and from flc_ivf_assign both equal the oracle's for the same centroids;
    neighbours at that n_probe equal the oracle's."""
    h = pipeline.HotPath(pipeline.Settings(exhaustive=False, n_probe=64))
    sp = synth.generate(44000, 29, mass_range=(1000.05, 1000.85))
    d = helpers.to_device(sp, h.device)
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order)
    ivf = h.build_ivf(v, b)
    nlist, cptr, bptr = _cpu(ivf.nlist), _cpu(ivf.centroid_ptr), _cpu(b.bucket_ptr)
    assert np.diff(bptr).max() > 20000 and nlist[: b.n_buckets].max() == 512 and ivf.max_nprobe == 64
    again = h.build_ivf(_drop_ell(v), b, centroids=ivf.centroids)  # flc_ivf_assign on the dense rows
    assert torch.equal(again.list_id, ivf.list_id) and torch.equal(again.probes, ivf.probes)
    cents, xs = _cpu(ivf.centroids), _cpu(v.x)
    shared = [cents[cptr[i]: cptr[i] + nlist[i]] if nlist[i] else None for i in range(b.n_buckets)]
    lid, probes = _cpu(ivf.list_id), _cpu(ivf.probes)
    for i in range(b.n_buckets):
        if nlist[i]:
            s, e = bptr[i], bptr[i + 1]
            p = oivf.n_probe_rule(int(nlist[i]), 64)
            assert np.array_equal(lid[s:e], oivf.assign_lists(xs[s:e], shared[i]))
            assert np.array_equal(probes[s:e, :p], oivf.probe_lists(xs[s:e], shared[i], p))
            assert (probes[s:e, p:] == -1).all()
    g = h.knn_graph(v, b, ivf)
    o = helpers.oracle_pipeline(sp, exhaustive=False, centroids=shared, vectors=xs, n_probe=64)
    ref = o["csr_cut"]
    assert np.array_equal(_cpu(g.indptr), ref.indptr) and np.array_equal(_cpu(g.indices), ref.indices)


def _train(n, seed, lo, hi, force_tiled=False, monkeypatch=None, want_bf16=False, tc_all=False, tc_dense=False,
           low_dim=400, negate=False):
    if monkeypatch is not None:
        if force_tiled:
            monkeypatch.setenv("FLC_KMEANS_FORCE_TILED", "1")
        else:
            monkeypatch.delenv("FLC_KMEANS_FORCE_TILED", raising=False)
        if tc_all:
            monkeypatch.delenv("FLC_KMEANS_SIMT_SMALL", raising=False)
        else:
            monkeypatch.setenv("FLC_KMEANS_SIMT_SMALL", "1")
        if tc_dense:
            monkeypatch.setenv("FLC_KMEANS_TC_DENSE", "1")
        else:
            monkeypatch.delenv("FLC_KMEANS_TC_DENSE", raising=False)
    h = pipeline.HotPath(pipeline.Settings(exhaustive=False, low_dim=low_dim))
    sp = helpers.dataset(n, seed, lo, hi)
    d = helpers.to_device(sp, h.device)
    if negate:  # every 7th peak negative: the operands of the assignment are no longer all >= 0
        d["intensity"] = torch.where(torch.arange(d["intensity"].shape[0], device=h.device) % 7 == 0,
                                     -d["intensity"], d["intensity"])
    b = h.bucket_sort(d["precursor_mz"], d["charge"])
    v = h.vectorize(d["mz"], d["intensity"], d["indptr"], b.order, want_bf16=want_bf16)
    return h, b, v, h.build_ivf(v, b)


@pytest.mark.parametrize("n,hi,bf16", [(6000, 1003.0, False), (30000, 1002.0, False), (30000, 1002.0, True),
                                        (12000, 1012.0, False), (24000, 1010.0, True), (36000, 1010.0, True)])
def test_kmeans_quality_close_to_oracle(n, hi, bf16):
    """n = 30000 over 2 Da makes buckets of ~7500 rows (128 lists), n = 24000 over 10 Da buckets
    of ~1200 rows (16 lists, 2 probes), n = 36000 over 10 Da buckets of ~1800 rows (32 lists, 4 probes):
    too large for the fused trainer, so those runs exercise the tiled trainer and its two-pass final
    assignment (float32 thread-per-row pass, float64 pass for the close calls and for > 32 lists); the other two are fused
    (two size classes).  With the bf16 rows the tiled trainer takes its assignment
    from the tensor cores.  Same arithmetic conventions as the oracle, so the
    objective is the same up to float32 arg-max near-ties."""
    h, b, v, ivf = _train(n, 29, 1000.0, hi, want_bf16=bf16)
    xs, bptr, nlist, cptr = _cpu(v.x), _cpu(b.bucket_ptr), _cpu(ivf.nlist), _cpu(ivf.centroid_ptr)
    cents = _cpu(ivf.centroids)
    lid = _cpu(ivf.list_id)
    checked = 0
    for i in range(b.n_buckets):
        if nlist[i] and checked < 3:
            xb_ = xs[bptr[i]: bptr[i + 1]]
            c_gpu = cents[cptr[i]: cptr[i] + nlist[i]]
            c_cpu = oivf.kmeans_train(xb_, int(nlist[i]))
            np.testing.assert_allclose(np.linalg.norm(c_gpu, axis=1), 1.0, atol=1e-5)
            obj_gpu = (xb_ @ c_gpu.T).max(axis=1).sum()
            obj_cpu = (xb_ @ c_cpu.T).max(axis=1).sum()
            assert obj_gpu >= 0.98 * obj_cpu
            # the final assignment and the probe lists belong to the trained centroids
            assert np.array_equal(lid[bptr[i]: bptr[i + 1]], oivf.assign_lists(xb_, c_gpu))
            p = oivf.n_probe_rule(int(nlist[i]), 32)
            assert np.array_equal(_cpu(ivf.probes)[bptr[i]: bptr[i + 1], :p], oivf.probe_lists(xb_, c_gpu, p))
            checked += 1
    assert checked > 0


def test_kmeans_fused_and_tiled_give_the_same_bits(monkeypatch):
    """List sums are fixed point, so the schedule (fused shared-memory trainer vs
    tiled multi-launch trainer with atomics, thread-per-row or tensor-core assignment)
    cannot change a single bit, and neither can a re-run."""
    _, b, _, fused = _train(9000, 31, 1000.0, 1010.0, False, monkeypatch)
    _, _, _, tiled = _train(9000, 31, 1000.0, 1010.0, True, monkeypatch)
    _, _, _, again = _train(9000, 31, 1000.0, 1010.0, True, monkeypatch)
    # with the bf16 rows the tiled trainer assigns on the tensor cores (close calls re-scored exactly);
    # "mixed" leaves the buckets of up to 32 lists (here: all) to the thread-per-row kernel
    # (sparse rows expanded in shared memory; "dense": the bf16 rows through TMA)
    _, _, _, tensor = _train(9000, 31, 1000.0, 1010.0, True, monkeypatch, want_bf16=True, tc_all=True)
    _, _, _, dense = _train(9000, 31, 1000.0, 1010.0, True, monkeypatch, want_bf16=True, tc_all=True, tc_dense=True)
    _, _, _, mixed = _train(9000, 31, 1000.0, 1010.0, True, monkeypatch, want_bf16=True)
    assert fused.total_centroids > 0
    for other in (tiled, again, tensor, dense, mixed):
        assert torch.equal(fused.centroids, other.centroids)
        assert torch.equal(fused.list_id, other.list_id)
        assert torch.equal(fused.probes, other.probes)


@pytest.mark.parametrize("low_dim,negate", [(200, False), (800, False), (400, True)])
def test_kmeans_tensor_core_schedules_other_shapes(monkeypatch, low_dim, negate):
    """low_dim = 800 is past the sparse-row kernel's operand buffer (dense bf16 rows through TMA instead);
    negative values switch the sparse-row kernel back to the general bf16 margin.  Same bits as the fused
    trainer either way."""
    kw = dict(low_dim=low_dim, negate=negate)
    _, _, _, fused = _train(9000, 33, 1000.0, 1010.0, False, monkeypatch, **kw)
    _, _, _, tensor = _train(9000, 33, 1000.0, 1010.0, True, monkeypatch, want_bf16=True, tc_all=True, **kw)
    assert fused.total_centroids > 0
    assert torch.equal(fused.centroids, tensor.centroids)
    assert torch.equal(fused.list_id, tensor.list_id)
    assert torch.equal(fused.probes, tensor.probes)


# ------------------------------------------------------------------ DBSCAN + split
def _random_graph(rng, n, max_deg):
    deg = rng.integers(1, max_deg + 1, n)
    indptr = np.r_[0, np.cumsum(deg)].astype(np.int64)
    indices = np.concatenate([np.r_[i, rng.integers(max(0, i - 40), min(n, i + 40), dd - 1)]
                              for i, dd in enumerate(deg)]).astype(np.int32)
    data = (rng.random(indices.shape[0]) * 0.2).astype(np.float32)
    data[indptr[:-1]] = 0
    return data, indices, indptr


@pytest.mark.parametrize("n,max_deg,seed", [(1, 1, 0), (50, 3, 1), (5000, 4, 2), (20000, 3, 3), (3000, 8, 4)])
def test_dbscan_random_directed_graphs(hp, n, max_deg, seed):
    data, indices, indptr = _random_graph(np.random.default_rng(seed), n, max_deg)
    dev = hp.device
    g = pipeline.KnnGraph(torch.from_numpy(data).to(dev), torch.from_numpy(indices).to(dev),
                          torch.from_numpy(indptr).to(dev), len(data), 0)
    labels, nc = hp.dbscan(g, n)
    ref = odb.dbscan_sklearn(data, indices, indptr, 0.1)
    assert np.array_equal(_cpu(labels), ref)  # labels, not just the partition
    assert nc == ref.max() + 1


def test_medoids_match_oracle(hp):
    """a16: representatives from the sparse matrix; random graphs with random
    labels exercise the quarter-of-the-cluster rule, ties and tiny clusters."""
    from oracle import medoids as omed

    for seed, n, max_deg, n_lab in [(0, 1, 1, 1), (1, 400, 6, 40), (2, 5000, 8, 300), (3, 3000, 12, 5)]:
        rng = np.random.default_rng(seed)
        data, indices, indptr = _random_graph(rng, n, max_deg)
        labels = rng.integers(-1, n_lab, n).astype(np.int32)
        nc = int(labels.max()) + 1
        dev = hp.device
        g = pipeline.KnnGraph(torch.from_numpy(data).to(dev), torch.from_numpy(indices).to(dev),
                              torch.from_numpy(indptr).to(dev), data.shape[0], 0)
        got = _cpu(hp.medoids(g, torch.from_numpy(labels).to(dev), nc)) if nc > 0 else np.zeros(0, np.int32)
        assert np.array_equal(got, omed.cluster_medoids(data, indices, indptr, labels))


def test_representatives_end_to_end():
    """Representatives follow the published rule on the FULL n_neighbors matrix: cluster members
    beyond eps count in the mean and in the quarter-of-the-cluster test.  Reference: the oracle's medoids on
    the oracle's uncut matrix.  Also in bucket ranges (max_pairs) and through the facade."""
    from falcon_b200.cluster import cluster as fcluster
    from oracle import medoids as omed

    h = pipeline.HotPath(pipeline.Settings(exhaustive=True, representatives=True, eps=0.2))
    sp = helpers.dataset(6000, 41, 1000.0, 1010.0)
    d = helpers.to_device(sp, h.device)
    labels, nc, keep = h.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    reps = _cpu(keep["representatives"])
    assert reps.shape == (nc,) and np.array_equal(_cpu(labels)[reps], np.arange(nc))
    o = helpers.oracle_pipeline(sp, exhaustive=True, eps=0.2)
    full = o["csr"]
    ref_rows = omed.cluster_medoids(full.data, full.indices, full.indptr, np.asarray(o["labels"]))
    assert odb.same_partition(_cpu(keep["sorted_labels"]), o["labels"])
    order = _cpu(keep["buckets"].order)
    assert sorted(reps.tolist()) == sorted(order[ref_rows].tolist())
    # the eps-cut matrix would pick other representatives for some clusters: the test can tell the two rules apart
    g = keep["graph"]
    cut_rows = omed.cluster_medoids(_cpu(g.dist), _cpu(g.indices), _cpu(g.indptr), _cpu(keep["sorted_labels"]))
    assert sorted(order[cut_rows].tolist()) != sorted(reps.tolist())
    # bucket ranges of at most ~2 buckets at a time give the same rows
    rows = h.representatives_exact(keep["vectors"], keep["buckets"], None, keep["sorted_labels"], nc, max_pairs=800_000)
    assert np.array_equal(order[_cpu(rows)], reps)
    # the facade with the published signature, fed the uncut matrix
    got = fcluster.get_cluster_representatives(np.asarray(o["labels"]), full.indptr, full.indices, full.data)
    assert np.array_equal(got, ref_rows)
    assert fcluster.get_cluster_representatives(np.full(3, -1), np.zeros(4, np.int64), np.zeros(0, np.int32),
                                                np.zeros(0, np.float32)) is None


def test_representatives_default_nprobe():
    """With the IVF index the uncut rows hold only members of probed lists; rule and ranges as above."""
    h = pipeline.HotPath(pipeline.Settings(representatives=True))
    sp = helpers.dataset(20000, 43, 1000.0, 1020.0)
    d = helpers.to_device(sp, h.device)
    labels, nc, keep = h.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], keep=True)
    reps = _cpu(keep["representatives"])
    assert reps.shape == (nc,) and np.array_equal(_cpu(labels)[reps], np.arange(nc))
    ivf = keep["ivf"]
    rows = h.representatives_exact(keep["vectors"], keep["buckets"], ivf, keep["sorted_labels"], nc, max_pairs=3_000_000)
    assert np.array_equal(_cpu(keep["buckets"].order)[_cpu(rows)], reps)
    # against the oracle's medoids on the uncut matrix the device builds (same index, no eps cut)
    from oracle import medoids as omed

    g = h.knn_graph(keep["vectors"], keep["buckets"], ivf, pair_capacity=int((np.diff(_cpu(keep["buckets"].bucket_ptr)) ** 2).sum()) + 1024,
                    eps_cut=False)
    ref = omed.cluster_medoids(_cpu(g.dist), _cpu(g.indices), _cpu(g.indptr), _cpu(keep["sorted_labels"]))
    assert np.array_equal(_cpu(rows), ref)


def test_dbscan_long_chain(hp):
    """A 30k-long one-way chain: pointer jumping must converge and give one cluster."""
    n = 30000
    indptr = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
    indices = np.stack([np.arange(n), np.minimum(np.arange(n) + 1, n - 1)], 1).reshape(-1).astype(np.int32)
    data = np.zeros(2 * n, np.float32)
    dev = hp.device
    g = pipeline.KnnGraph(torch.from_numpy(data).to(dev), torch.from_numpy(indices).to(dev),
                          torch.from_numpy(indptr).to(dev), 2 * n, 0)
    labels, nc = hp.dbscan(g, n)
    assert nc == 1 and (_cpu(labels) == 0).all()


@pytest.mark.parametrize("mode,tol", [("ppm", 20.0), ("Da", 0.02)])
@pytest.mark.parametrize("values_sorted", [False, True])
def test_split_clusters(mode, tol, values_sorted):
    h = pipeline.HotPath(pipeline.Settings(precursor_tol_mass=tol, precursor_tol_mode=mode))
    rng = np.random.default_rng(31)
    n = 30000
    labels = rng.integers(-1, 4000, n).astype(np.int32)
    labels[rng.random(n) < 0.2] = -1
    big = rng.random(n) < 0.1
    labels[big] = 4000 + rng.integers(0, 3, big.sum())  # a few clusters with ~1000 members
    centre = 400.0 + (np.maximum(labels, 0) % 997)
    mz = centre + rng.normal(0, 1.0, n) * (centre * 12e-6 if mode == "ppm" else 0.012)
    dup = rng.random(n) < 0.05
    mz[dup] = centre[dup]  # exact ties
    if values_sorted:
        order = np.argsort(mz, kind="stable")
        labels, mz = labels[order], mz[order]
    dev = h.device
    out, nc = h.split(torch.from_numpy(labels).to(dev), torch.from_numpy(mz).to(dev), values_sorted)
    out = _cpu(out)
    ref = np.full(n, -1, np.int64)
    nxt = 0
    for l in np.unique(labels[labels >= 0]):
        members = np.flatnonzero(labels == l)
        sub, k = odb.postprocess_cluster(mz[members], None, tol, mode, None)
        ref[members[sub >= 0]] = sub[sub >= 0] + nxt
        nxt += k
    assert nc == nxt
    assert odb.same_partition(out, ref)
    assert sorted(np.unique(out[out >= 0]).tolist()) == list(range(nc))


def test_split_golden_cases(hp, golden_dir):
    """The reference's own _postprocess_cluster outputs (tests/golden/postprocess.npz), with and
    without the retention-time cut (cluster.py:418-429)."""
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    dev = hp.device
    n_checked = n_rt = 0
    for i in range(int(g["n_post"])):
        rt_tol = float(g[f"post_rttol{i}"])
        v, mode, tol = g[f"post_v{i}"], str(g[f"post_mode{i}"]), float(g[f"post_tol{i}"])
        h = pipeline.HotPath(pipeline.Settings(precursor_tol_mass=tol, precursor_tol_mode=mode,
                                               rt_tol=rt_tol if rt_tol >= 0 else None))
        rt = torch.from_numpy(g[f"post_rt{i}"]).to(dev) if rt_tol >= 0 else None
        out, nc = h.split(torch.zeros(len(v), dtype=torch.int32, device=dev), torch.from_numpy(v).to(dev), False, rt=rt)
        assert nc == int(g[f"post_k{i}"]), i
        assert odb.same_partition(_cpu(out), g[f"post_labels{i}"]), i
        n_checked += 1
        n_rt += rt_tol >= 0
    assert n_checked >= 120 and n_rt >= 60


def test_split_golden_cases_in_one_call(hp, golden_dir):
    """All golden retention-time cases as the DBSCAN clusters of ONE call (plus noise rows)."""
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    cases = [i for i in range(int(g["n_post"])) if float(g[f"post_rttol{i}"]) >= 0 and str(g[f"post_mode{i}"]) == "ppm"]
    labels, mz, rt, want = [], [], [], []
    nxt = 0
    for c, i in enumerate(cases):
        v = g[f"post_v{i}"]
        labels.append(np.full(len(v), c))
        mz.append(v)
        rt.append(g[f"post_rt{i}"])
        w = g[f"post_labels{i}"].astype(np.int64)
        want.append(np.where(w >= 0, w - 7 + nxt, -1))
        nxt += int(g[f"post_k{i}"])
    labels.append(np.full(5, -1)); mz.append(np.full(5, 500.0)); rt.append(np.zeros(5)); want.append(np.full(5, -1))
    labels, mz, rt, want = (np.concatenate(a) for a in (labels, mz, rt, want))
    perm = np.random.default_rng(5).permutation(len(labels))
    labels, mz, rt, want = labels[perm], mz[perm], rt[perm], want[perm]
    h = pipeline.HotPath(pipeline.Settings(precursor_tol_mass=20.0, precursor_tol_mode="ppm", rt_tol=5.0))
    dev = h.device
    out, nc = h.split(torch.from_numpy(labels.astype(np.int32)).to(dev), torch.from_numpy(mz).to(dev), False,
                      rt=torch.from_numpy(rt).to(dev))
    assert nc == nxt
    assert odb.same_partition(_cpu(out), want)
    assert sorted(np.unique(_cpu(out)[_cpu(out) >= 0]).tolist()) == list(range(nc))


@pytest.mark.parametrize("mode,tol,rt_tol", [("ppm", 20.0, 3.0), ("Da", 0.02, 0.5), ("ppm", 10.0, 40.0)])
def test_split_clusters_with_rt(mode, tol, rt_tol):
    """Random DBSCAN clusters (a few with ~1000 members, exact ties in both columns) against the
    oracle's postprocess_cluster, which cuts the restated linkage with scipy's fcluster itself."""
    h = pipeline.HotPath(pipeline.Settings(precursor_tol_mass=tol, precursor_tol_mode=mode, rt_tol=rt_tol))
    rng = np.random.default_rng(77)
    n = 20000
    labels = rng.integers(-1, 3000, n).astype(np.int32)
    labels[rng.random(n) < 0.2] = -1
    big = rng.random(n) < 0.1
    labels[big] = 3000 + rng.integers(0, 2, big.sum())
    centre = 400.0 + (np.maximum(labels, 0) % 997)
    mz = centre + rng.normal(0, 1.0, n) * (centre * 12e-6 if mode == "ppm" else 0.012)
    dup = rng.random(n) < 0.05
    mz[dup] = centre[dup]
    rt = rng.uniform(0, 8 * rt_tol, n)
    rt[labels >= 3000] = rng.uniform(0, 500 * rt_tol, int((labels >= 3000).sum()))
    rt[rng.random(n) < 0.05] = 2 * rt_tol  # ties
    if rng.random() < 0.5:
        rt = rt.astype(np.float32).astype(np.float64)
    dev = h.device
    out, nc = h.split(torch.from_numpy(labels).to(dev), torch.from_numpy(mz).to(dev), False,
                      rt=torch.from_numpy(rt).to(dev))
    out = _cpu(out)
    ref = np.full(n, -1, np.int64)
    nxt = 0
    for l in np.unique(labels[labels >= 0]):
        members = np.flatnonzero(labels == l)
        sub, k = odb.postprocess_cluster(mz[members], rt[members], tol, mode, rt_tol)
        ref[members[sub >= 0]] = sub[sub >= 0] + nxt
        nxt += k
    assert nc == nxt
    assert odb.same_partition(out, ref)
    assert sorted(np.unique(out[out >= 0]).tolist()) == list(range(nc))


def test_rt_tolerance_needs_retention_times(hp):
    h = pipeline.HotPath(pipeline.Settings(rt_tol=5.0))
    z = torch.zeros(4, dtype=torch.int32, device=h.device)
    with pytest.raises(ValueError, match="retention times"):
        h.split(z, torch.ones(4, dtype=torch.float64, device=h.device), False)


# ------------------------------------------------------------------ preprocessing
def _raw_spectra(n, seed):
    """Unprocessed-looking spectra: 0-400 peaks over 50-2000 m/z (ascending), intensities with
    exact ties and zeros, a peak planted on the precursor m/z of every charge state."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 400, n)
    counts[rng.random(n) < 0.05] = rng.integers(0, 8, int((rng.random(n) < 0.05).sum()) or 1)[0]
    pmz = rng.uniform(300.0, 1200.0, n)
    z = rng.integers(0, 5, n).astype(np.int32)
    mzs, ints, indptr = [], [], [0]
    for i in range(n):
        m = rng.uniform(50.0, 2000.0, counts[i])
        zz = max(int(z[i]), 1)
        neutral = (pmz[i] - 1.0072766) * zz
        planted = [neutral / c + 1.0072766 + rng.uniform(-2.0, 2.0) for c in range(1, zz + 1)]
        m = np.sort(np.r_[m, planted]).astype(np.float32)
        it = np.round(rng.lognormal(3.0, 1.5, m.shape[0]), 0 if i % 3 == 0 else 3).astype(np.float32)
        it[rng.random(m.shape[0]) < 0.02] = 0.0
        mzs.append(m)
        ints.append(it)
        indptr.append(indptr[-1] + m.shape[0])
    return synth.SpectrumSet(np.concatenate(mzs), np.concatenate(ints), np.asarray(indptr, np.int64), pmz, z,
                             np.zeros(n, np.float32))


@pytest.mark.parametrize("kw", [
    dict(min_peaks=5, min_mz_range=250.0, mz_min=101.0, mz_max=1500.0, remove_precursor_tolerance=1.5,
         min_intensity=0.01, max_peaks_used=50, scaling=None),  # falcon's defaults (config.py:127-183)
    dict(min_peaks=5, min_mz_range=250.0, mz_min=101.0, mz_max=1500.0, remove_precursor_tolerance=1.5,
         min_intensity=0.01, max_peaks_used=50, scaling="rank"),
    dict(min_peaks=3, min_mz_range=100.0, mz_min=None, mz_max=None, remove_precursor_tolerance=None,
         min_intensity=None, max_peaks_used=None, scaling="root"),
    dict(min_peaks=12, min_mz_range=0.0, mz_min=200.0, mz_max=None, remove_precursor_tolerance=0.5,
         min_intensity=0.2, max_peaks_used=None, scaling="log"),
    dict(min_peaks=5, min_mz_range=600.0, mz_min=None, mz_max=900.0, remove_precursor_tolerance=None,
         min_intensity=None, max_peaks_used=7, scaling=None),
])
def test_preprocess_matches_oracle(hp, kw):
    from oracle import preprocess as opre

    sp = _raw_spectra(1500, 11)
    dev = hp.device
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    mz, it, indptr, valid = hp.preprocess(t(sp.mz), t(sp.intensity), t(sp.indptr), t(sp.precursor_mz),
                                          t(sp.precursor_charge), **{**kw, "scaling": kw["scaling"] or "off"})
    v_ref, mz_ref, it_ref, ip_ref = opre.process_spectra(sp, **kw)
    assert 0 < v_ref.sum() < len(sp)  # both outcomes occur
    assert np.array_equal(_cpu(valid).astype(bool), v_ref)
    assert np.array_equal(_cpu(indptr), ip_ref)
    assert np.array_equal(_cpu(mz), mz_ref)
    np.testing.assert_allclose(_cpu(it), it_ref, rtol=0, atol=1e-6)


def test_preprocess_edge_cases(hp):
    dev = hp.device
    e = torch.empty(0, device=dev)
    mz, it, indptr, valid = hp.preprocess(e.float(), e.float(), torch.zeros(1, dtype=torch.int64, device=dev),
                                          e.double(), e.int())
    assert mz.shape == (0,) and valid.shape == (0,) and _cpu(indptr).tolist() == [0]
    with pytest.raises(ValueError):
        hp.preprocess(e.float(), e.float(), torch.zeros(1, dtype=torch.int64, device=dev), e.double(), e.int(),
                      scaling="cube")
