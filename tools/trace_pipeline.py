#!/usr/bin/env python
"""Timeline of the software-pipelined host path: per batch, when its H2D copies and
its kernels start and end (CUDA events, ms since the first event)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_b200 import pipeline, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
sp = synth.generate(n, 42)
hp = pipeline.HotPath(pipeline.Settings())
host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in dict(
    mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr, precursor_mz=sp.precursor_mz,
    charge=sp.precursor_charge).items()}
out = torch.empty(n, dtype=torch.int32).pin_memory()
mp = int(np.diff(sp.indptr).max())


def stage():
    t0 = time.perf_counter()
    st = hp.stage_host(host["mz"], host["intensity"], host["indptr"], host["precursor_mz"], host["charge"],
                       max_peaks=mp)
    st["host_ms"] = (time.perf_counter() - t0) * 1e3
    return st


for rep in range(2):
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    rows = []
    staged = stage()
    h0 = time.perf_counter()
    for i in range(K):
        nxt = stage() if i + 1 < K else None
        a = torch.cuda.Event(enable_timing=True)
        a.record()
        t0 = time.perf_counter()
        hp.run_staged(staged, labels_out=out)
        host_run = (time.perf_counter() - t0) * 1e3
        b = torch.cuda.Event(enable_timing=True)
        b.record()
        cp = torch.cuda.Event(enable_timing=True)
        cp.record(hp._copy_stream)
        rows.append((a, b, cp, staged["host_ms"], host_run, (time.perf_counter() - h0) * 1e3))
        staged = nxt
    torch.cuda.synchronize()
    if rep == 1:
        for i, (a, b, cp, hs, hr, hw) in enumerate(rows):
            print(f"batch {i}: compute-stream enq {base.elapsed_time(a):7.2f} end {base.elapsed_time(b):7.2f}  "
                  f"copy-stream idle at {base.elapsed_time(cp):7.2f}  host: stage {hs:5.2f} run {hr:5.2f} wall {hw:7.2f}")
