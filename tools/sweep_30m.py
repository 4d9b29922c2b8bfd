#!/usr/bin/env python
"""BASELINE configs[4]: whole-box sweep of low_dim 200/400/800 x eps 0.05-0.30 over ONE data set of 30 M
synthetic spectra (240 bucket-aligned chunks dealt to the ranks: whole precursor buckets per GPU).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_30m.py [--total 30000000]

Per setting: the resident step (labels gathered over NVLink peer memory inside the step), max over ranks
of the CUDA-event time, clusters / noise summed over ranks, peak HBM reserved on any rank.  Rank 0 prints one
JSON document (kept in profiles/)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from falcon_b200 import distributed as fdist, pipeline, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=30_000_000)
    ap.add_argument("--chunks", type=int, default=240)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--low-dims", type=int, nargs="+", default=[200, 400, 800])
    ap.add_argument("--eps", type=float, nargs="+", default=[0.05, 0.10, 0.20, 0.30])
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = args.total // args.chunks * args.chunks
    mine = range(args.chunks * rank // world, args.chunks * (rank + 1) // world)
    workers = max(1, min(16, len(os.sched_getaffinity(0)) // world))
    t0 = time.perf_counter()
    sp = synth.generate_chunks(total, args.chunks, mine, workers=workers)
    gen_s = time.perf_counter() - t0
    rows = []
    wl = None
    for low_dim in args.low_dims:
        for eps in args.eps:
            del wl
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            hp = pipeline.HotPath(pipeline.Settings(low_dim=low_dim, eps=eps), dev)
            wl = bench.Workload(torch, fdist, hp, sp, dev, world)
            for _ in range(2):  # warm-up: the first step learns the sizes, the second runs without read-backs
                wl.step_resident()
            wl.wait_gathers()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                out, nc = wl.step_resident()
            wl.wait_gathers()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.steps
            lab = out if world == 1 else out[rank]
            agg = torch.tensor([ms, float(nc), float((lab[: len(sp)] < 0).sum().item()), float(len(sp)),
                                float(torch.cuda.max_memory_reserved() / 2 ** 30), float(getattr(hp, "spec_misses", 0))],
                               dtype=torch.float64, device=dev)
            mx = agg.clone()
            if world > 1:
                dist.all_reduce(agg, op=dist.ReduceOp.SUM)
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            if rank == 0:
                rows.append({"low_dim": low_dim, "eps": eps, "ms_per_step": mx[0].item(),
                             "spectra_per_s": agg[3].item() / (mx[0].item() * 1e-3), "n_clusters": int(agg[1].item()),
                             "noise": int(agg[2].item()), "clustered_fraction": 1.0 - agg[2].item() / agg[3].item(),
                             "hbm_reserved_gib_max_rank": mx[4].item(), "sync_free_redos": int(agg[5].item())})
                print(f"[sweep] {rows[-1]}", file=sys.stderr, flush=True)
    if rank == 0:
        print(json.dumps({"config": f"configs[4]: {total} synthetic spectra, {args.chunks} bucket-aligned chunks over "
                                    f"{world} GPU(s), falcon defaults except low_dim / eps", "n_gpus": world,
                          "total_spectra": total, "steps": args.steps, "generation_s_rank0": gen_s,
                          "label_gather": wl.gather_mode, "rows": rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
