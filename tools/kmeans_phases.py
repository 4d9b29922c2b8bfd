#!/usr/bin/env python
"""Per-phase cycle shares of the fused k-means trainer (FLC_KMEANS_TIMING=1)."""
import ctypes as C
import os
import sys

os.environ["FLC_KMEANS_TIMING"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_b200 import _lib, pipeline, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sp = synth.generate(n, 42)
hp = pipeline.HotPath(pipeline.Settings())
dev = hp.device
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr,
                                                     precursor_mz=sp.precursor_mz, charge=sp.precursor_charge).items()}
buf = (C.c_ulonglong * 16)()
for i in range(3):
    hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
    _lib.check(_lib.lib.flc_debug_kmeans_timing(buf))
names = ["queue+nnz", "row load", "init", "assign", "update", "means", "normalise", "final+out"]
tot = sum(buf[i] for i in range(8))
print(f"buckets {buf[9]}, iterations/bucket {buf[8] / max(buf[9], 1):.2f}, cycles/bucket {tot / max(buf[9], 1):.0f}")
print(f"slowest bucket: {buf[10]} cycles, rows {buf[11] >> 32}, lists {(buf[11] >> 8) & 0xffffff}")
for i, nm in enumerate(names):
    print(f"  {nm:10s} {buf[i] / max(buf[9], 1):9.0f} cycles/bucket {100 * buf[i] / max(tot, 1):5.1f}%")
