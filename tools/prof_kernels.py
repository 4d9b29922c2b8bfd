import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_b200 import _lib, pipeline, synth
sp = synth.generate(1_000_000, 42)
hp = pipeline.HotPath(pipeline.Settings())
dev = hp.device
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(mz=sp.mz, intensity=sp.intensity, indptr=sp.indptr, precursor_mz=sp.precursor_mz, charge=sp.precursor_charge).items()}
for i in range(3):
    hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
_lib.profile_reset(); _lib.profile_enable(True)
for i in range(3):
    hp.run(d["mz"], d["intensity"], d["indptr"], d["precursor_mz"], d["charge"])
torch.cuda.synchronize()
_lib.profile_enable(False)
for k, (ms, n) in _lib.profile_summary().items():
    if "kmeans" in k: print(k, round(ms / 3, 4), n)
