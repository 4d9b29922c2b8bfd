"""CPU-only checks: the C-ABI library loads, exports every symbol declared in
include/falcon_b200.h, host-side functions agree with the oracle, and the
product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "falcon_b200.h")).read()
    return sorted(set(re.findall(r"FLC_API\s+[\w\s\*]+?\b(flc_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from falcon_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/falcon_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == syms


def test_library_has_no_cuda_driver_link_dependency():
    from falcon_b200 import _lib

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "libtorch" not in out


def test_get_dim_host_matches_reference_golden(golden_dir):
    import json

    from falcon_b200.cluster import spectrum

    for row in json.load(open(os.path.join(golden_dir, "get_dim.json"))):
        assert spectrum.get_dim(row["min_mz"], row["max_mz"], row["bin_size"]) == (
            row["vec_len"], row["start"], row["end"])


def test_invalid_arguments_raise_value_error():
    from falcon_b200 import _lib

    n, s, e = ctypes.c_uint32(), ctypes.c_float(), ctypes.c_float()
    rc = _lib.lib.flc_get_dim(101.0, 1500.0, 0.0, ctypes.byref(n), ctypes.byref(s), ctypes.byref(e))
    with pytest.raises(ValueError, match="bin_size"):
        _lib.check(rc)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "falcon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
    code = "import sys; import falcon_b200.cluster, falcon_b200.pipeline; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_no_cuda_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from falcon_b200 import pipeline, synth
    from falcon_b200.cluster import spectrum

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pipeline.HotPath()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        spectrum.to_vector_parallel(synth.generate(4), 400, 100.95, 1500.0, 0.05)


def test_synthetic_generator_contract():
    from falcon_b200 import synth

    s = synth.generate(3000, 7)
    assert len(s) == 3000 and s.indptr[-1] == s.mz.shape[0]
    cnt = np.diff(s.indptr)
    assert cnt.min() >= 1 and cnt.max() <= 50
    assert s.mz.dtype == np.float32 and s.intensity.dtype == np.float32
    assert s.mz.min() >= 101.0 and s.mz.max() <= 1500.0
    # peaks sorted by m/z inside every spectrum, unit L2 intensity
    row = np.repeat(np.arange(3000), cnt)
    assert (np.diff(s.mz)[row[1:] == row[:-1]] >= 0).all()
    ss = np.zeros(3000)
    np.add.at(ss, row, s.intensity.astype(np.float64) ** 2)
    np.testing.assert_allclose(ss, 1.0, atol=1e-5)
    assert set(np.unique(s.precursor_charge)) <= {2, 3}
    s2 = synth.generate(3000, 7)
    assert np.array_equal(s.mz, s2.mz) and np.array_equal(s.precursor_mz, s2.precursor_mz)
    t = s.take(np.array([5, 1, 7]))
    assert np.array_equal(t.mz[: cnt[5]], s.mz[s.indptr[5]: s.indptr[6]])
    d = s.take(np.arange(10)).as_dicts()
    r = synth.SpectrumSet.from_dicts(d)
    assert np.array_equal(r.mz, s.take(np.arange(10)).mz)
