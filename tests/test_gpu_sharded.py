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
