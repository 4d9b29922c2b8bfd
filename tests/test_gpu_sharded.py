"""One data set over several GPUs: ``distributed.cluster_sharded`` against the single-GPU
run -- whole buckets per rank, oversized buckets cut with a tolerance halo, labels and representatives
gathered.  The NCCL test needs two visible GPUs (``gpurun --gpus 2``); the one-process tests exercise
the same planning / halo / row-gather code on one GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from falcon_b200 import distributed as fd  # noqa: E402
from falcon_b200 import pipeline, synth  # noqa: E402
from tests import helpers  # noqa: E402


def _single(sp, settings):
    labels, nc, _ = pipeline.cluster_host(sp, settings)
    return labels, nc


@pytest.mark.parametrize("exhaustive,cap", [(True, 700), (True, None), (False, None)])
def test_sharded_one_rank_equals_plain_run(exhaustive, cap):
    """Buckets of ~1 700 rows cut at 700 rows: three pieces per bucket with halos, rows re-assembled,
    DBSCAN per bucket -- the partition of the plain run.  Without a cap: the plain path through the planner."""
    sp = helpers.dataset(20000, 46, 1000.0, 1012.0)
    s = pipeline.Settings(exhaustive=exhaustive, representatives=True)
    ref, nc_ref = _single(sp, s)
    labels, nc, reps = fd.cluster_sharded(sp, s, bucket_cap=cap)
    assert nc == nc_ref and fd.same_partition(labels, ref)
    assert reps.shape[0] == nc and (labels[reps] == np.arange(nc)).all()


def test_sharded_one_rank_rt_and_da():
    sp = helpers.dataset(12000, 47, 1000.0, 1008.0)
    s = pipeline.Settings(exhaustive=True, rt_tol=15.0, precursor_tol_mass=0.02, precursor_tol_mode="Da")
    ref, nc_ref = _single(sp, s)
    labels, nc, reps = fd.cluster_sharded(sp, s, bucket_cap=500)
    assert reps is None and nc == nc_ref and fd.same_partition(labels, ref)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    # 60 % of the spectra in a 12-Da window (buckets of ~5 000 rows, cut at 2 000), the rest spread out
    a = synth.generate(int(n * 0.6), 50, mass_range=(1000.0, 1012.0))
    b = synth.generate(n - len(a), 51)
    sp = synth.concat([a, b])
    out = {}
    for exhaustive, cap in ((True, 2000), (False, None)):
        s = pipeline.Settings(exhaustive=exhaustive, representatives=True)
        labels, nc, reps = fd.cluster_sharded(sp, s, device=rank, bucket_cap=cap)
        if rank == 0:
            ref, nc_ref, _ = pipeline.cluster_host(sp, s, device=0)
            out[exhaustive] = (nc == nc_ref, fd.same_partition(labels, ref),
                               bool(reps.shape[0] == nc and (labels[reps] == np.arange(nc)).all()), nc)
        dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_two_ranks_nccl_equal_single_gpu():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 200_000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for exhaustive in (True, False):
        same_nc, same, reps_ok, nc = res[0][exhaustive]
        assert same_nc and same and reps_ok and nc > 1000


def _peer_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n_max = 30000
    ok = []
    hp = pipeline.HotPath(pipeline.Settings(), dev)
    plain = pipeline.HotPath(pipeline.Settings(), dev)
    sink = fd.PeerLabelGather(n_max, dev)
    hp.label_sink = sink
    for batch in range(5):  # more batches than slot generations; a different length on every rank and batch
        n = n_max - 1000 * rank - 777 * batch
        sp = synth.generate(n, 60 + 10 * batch + rank, mass_range=(1000.0, 1030.0))
        d = helpers.to_device(sp, dev)
        labels, nc = hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
        ref, nc_ref = plain.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
        want, want_lens = fd.gather_labels_padded(ref, nc_ref, max_len=n_max)  # the NCCL gather
        sink.wait()
        got, lens = sink.last
        ok.append(bool(nc == nc_ref and torch.equal(labels, ref) and torch.equal(got, want)
                       and torch.equal(lens, want_lens) and int(got.max()) + 1 > nc))
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_peer_memory_label_gather_equals_nccl_gather():
    """flc_scatter_labels_peers + barrier + flc_relabel_gathered (NVLink peer stores from the step's last kernel)
    give exactly the labels of the NCCL all-gather, batch after batch (slot reuse), ragged lengths."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res[0] == [True] * 5 and res[1] == [True] * 5


def test_relabel_gathered_single_process():
    """The relabel kernel on hand-made slots (any GPU count): offsets from the headers, -1 kept, padding -1."""
    import ctypes as C

    from falcon_b200._lib import check, lib, ptr

    dev = torch.device("cuda", 0)
    max_len = 8
    raw = [[5, 2, 0, 0, 0, -1, 1, 1, 0, 99, 99, 99], [3, 1, 0, 0, -1, 0, 0, 77, 77, 77, 77, 77],
           [0, 0, 0, 0, 9, 9, 9, 9, 9, 9, 9, 9]]
    slots = [torch.tensor(r, dtype=torch.int32, device=dev) for r in raw]
    out = torch.empty((3, max_len), dtype=torch.int32, device=dev)
    lens = torch.empty(3, dtype=torch.int64, device=dev)
    ptrs = (C.c_void_p * 3)(*[t.data_ptr() for t in slots])
    check(lib.flc_relabel_gathered(ptrs, 3, max_len, ptr(out), ptr(lens), None))
    assert out.cpu().tolist() == [[0, -1, 1, 1, 0, -1, -1, -1], [-1, 2, 2, -1, -1, -1, -1, -1], [-1] * 8]
    assert lens.cpu().tolist() == [5, 3, 0]
    # the scatter kernel writes header + labels in input order into the slot
    buf = torch.full((12,), -7, dtype=torch.int32, device=dev)
    own = (C.c_void_p * 1)(buf.data_ptr())
    lab = torch.tensor([4, -1, 2], dtype=torch.int32, device=dev)
    order = torch.tensor([2, 0, 1], dtype=torch.int32, device=dev)
    check(lib.flc_scatter_labels_peers(ptr(lab), ptr(order), 3, None, 5, own, 1, 1, None))
    assert buf.cpu().tolist() == [-7, 3, 5, 0, 0, -1, 2, 4, -7, -7, -7, -7]


def _cli_worker(rank, world, port, argv, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from falcon_b200 import falcon as fmain

    rc = fmain.main(list(argv))
    q.put((rank, rc))
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_falcon_command_on_two_gpus_equals_one_gpu(tmp_path):
    """The `falcon` command under a 2-rank launch (what torchrun sets up): every charge clustered by both GPUs
    through cluster_sharded, rank 0 writes.  Same partition, same number of representatives as the 1-GPU run."""
    import pandas as pd
    import torch.multiprocessing as mp

    from falcon_b200 import falcon as fmain
    from falcon_b200.ms_io import mgf_io

    sp = synth.generate(6000, 19, mass_range=(1000.0, 1010.0))
    path = str(tmp_path / "in.mgf")
    mgf_io.write_spectra(path, sp.as_dicts())
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    assert fmain.main([path, one, "--exhaustive", "--export_representatives"]) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cli_worker, args=(r, 2, port, [path, two, "--exhaustive", "--export_representatives"], q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res == {0: 0, 1: 0}
    a = pd.read_csv(one + ".csv", comment="#")
    b = pd.read_csv(two + ".csv", comment="#")
    assert a["spectrum_id"].tolist() == b["spectrum_id"].tolist()
    assert fd.same_partition(a["cluster"].to_numpy(), b["cluster"].to_numpy())
    ra, rb = list(mgf_io.get_spectra(one + ".mgf")), list(mgf_io.get_spectra(two + ".mgf"))
    assert len(ra) == len(rb) == a["cluster"][a["cluster"] >= 0].nunique()
    assert sorted(r["identifier"] for r in ra) == sorted(r["identifier"] for r in rb)
