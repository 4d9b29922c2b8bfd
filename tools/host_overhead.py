#!/usr/bin/env python
"""Host wall time vs device time of every stage call of one resident step."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_b200 import pipeline, synth  # noqa: E402

sp = synth.generate(1_000_000, 42)
hp = pipeline.HotPath(pipeline.Settings(), profile=True)
dev = hp.device
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr,
                                                     precursor_mz=sp.precursor_mz, charge=sp.precursor_charge).items()}
mp = int((sp.indptr[1:] - sp.indptr[:-1]).max())
wall = {}
orig = {}
for name in ("bucket_sort", "vectorize", "build_ivf", "knn_graph", "dbscan", "split"):
    fn = getattr(hp, name)
    orig[name] = fn

    def wrap(*a, _fn=fn, _name=name, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = _fn(*a, **k)
        torch.cuda.synchronize()
        wall[_name] = wall.get(_name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    setattr(hp, name, wrap)
for i in range(6):
    if i == 3:
        wall.clear()
        hp.timer.events.clear()
    hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"], max_peaks=mp)
gpu = hp.timer.result()
print("stage        wall(ms)  device(ms)  (3 steps averaged; wall includes a sync on both sides)")
alias = {"build_ivf": "ivf_train", "knn_graph": None}
for name in wall:
    if name == "knn_graph":
        g = gpu.get("scan", 0) + gpu.get("knn_csr", 0)
    else:
        g = gpu.get(alias.get(name, name), 0)
    print(f"{name:12s} {wall[name] / 3:8.3f} {g / 3:10.3f}")
