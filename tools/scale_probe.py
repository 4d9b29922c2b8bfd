"""Where does the extra step time at N >= 2 come from?  torchrun --nproc-per-node N tools/scale_probe.py
Times the resident step on every rank with the label gather (a) off, (b) on the compute stream, (c) on the
side stream (bench.py's way), and the gather alone; prints per-rank numbers (no max-reduction) so that a slow
rank shows."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from falcon_b200 import distributed as fdist, pipeline, synth  # noqa: E402


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("PROBE_N", "1000000"))
    sp = synth.generate(n, 42 + rank)
    hp = pipeline.HotPath(pipeline.Settings(), dev)  # plain: no label sink
    wl = bench.Workload(torch, fdist, pipeline.HotPath(pipeline.Settings(), dev), sp, dev, world)

    def run(fn, steps=10, warm=3, barrier_each=False):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            fn()
        wl.wait_gathers()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, (time.perf_counter() - t0) * 1e3 / steps

    d = wl.d

    def no_gather():
        return hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], max_peaks=wl.max_peaks)

    def main_stream_gather():
        labels, nc = no_gather()
        if world > 1:
            fdist.gather_labels_padded(labels, nc, max_len=n)

    lab, nc = no_gather()

    def gather_only():
        if world > 1:
            fdist.gather_labels_padded(lab, nc, max_len=n)

    res = {"no_gather": run(no_gather), "nccl_main_stream": run(main_stream_gather),
           f"bench_step[{wl.gather_mode.split()[0]}]": run(wl.step_resident), "nccl_gather_only": run(gather_only)}
    for r in range(world):
        if r == rank:
            print(f"rank {rank}/{world} n={n}: " + "  ".join(f"{k} {v[0]:.3f} ms (host {v[1]:.3f})" for k, v in res.items()),
                  flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
