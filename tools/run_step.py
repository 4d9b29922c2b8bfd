#!/usr/bin/env python
"""Run the hot path a few times on synthetic spectra (for ncu / compute-sanitizer).

    python tools/run_step.py [--n 1000000] [--steps 3] [--exhaustive] [--mass-range LO HI]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_b200 import pipeline, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--exhaustive", action="store_true")
ap.add_argument("--mass-range", type=float, nargs=2, default=None)
args = ap.parse_args()
kw = {"mass_range": tuple(args.mass_range)} if args.mass_range else {}
sp = synth.generate(args.n, 42, **kw)
hp = pipeline.HotPath(pipeline.Settings(exhaustive=args.exhaustive))
dev = hp.device
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(
    mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr, precursor_mz=sp.precursor_mz,
    charge=sp.precursor_charge).items()}
import time  # noqa: E402

for i in range(args.steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    labels, nc = hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
    b.record()
    torch.cuda.synchronize()
    print(f"step {i}: {nc} clusters, {a.elapsed_time(b):.3f} ms on the stream, {(time.perf_counter() - t0) * 1e3:.3f} ms wall, "
          f"reserved {torch.cuda.memory_reserved() >> 20} MiB", flush=True)
