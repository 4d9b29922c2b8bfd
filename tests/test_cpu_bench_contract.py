"""bench.py's contract, as far as it can be checked without a GPU: the reference arm's JSON
line, the roofline arithmetic, the clock sampler's failure modes, and that the main arm refuses
to run without a CUDA device (the product has no CPU fallback)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "20000",
                          "--cpu-sample", "4000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "spectra_per_sec_end_to_end"
    assert line["unit"] == "spectra/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_is_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "20000",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_main_arm_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", "1000", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def test_kernel_rooflines_arithmetic(bench):
    n, p, d, ldb, w = 1000, 30000, 400, 400, 56
    stats = {"n": n, "n_peaks": p, "low_dim": d, "ld_bf16": ldb, "dense_f32": 0, "ell_width": w, "n_pairs": 3000,
             "nnz": 2500, "steps": 2, "required_pairs": 1.0e6, "computed_pairs": 6.0e5, "ivf_rows": 800,
             "ivf_nnz": 20000, "total_centroids": 16, "max_nprobe": 2,
             "kernels": {"vectorize": (2.0, 2), "scan_tc": (4.0, 2), "kmeans_tc_sparse": (1.0, 20)}}
    peaks = {"hbm_gbs": 6000.0, "bf16_tflops": 1500.0, "_source": "test"}
    r = bench.kernel_rooflines(stats, peaks)
    vec_bytes = p * 8 + (n + 1) * 8 + n * 4 + n * ldb * 2 + n * (w * 6 + 2)
    assert r["vectorize"]["algorithmic_bytes_per_launch"] == vec_bytes
    assert r["vectorize"]["achieved"] == pytest.approx(vec_bytes / 1e-3 / 1e9)
    assert r["vectorize"]["frac"] == pytest.approx(r["vectorize"]["achieved"] / 6000.0)
    # the survey's formula also counts the dense float32 row
    assert r["vectorize"]["survey_8d"]["algorithmic_bytes_per_launch"] == p * 8 + (n + 1) * 8 + n * d * 4 + n * ldb * 2
    t = r["scan_tc"]["tensor"]
    # the tensor roofline is over the executed FLOPs (a fraction of peak <= 1 by construction); the pairs the
    # IVF semantics require are reported as a rate beside it
    assert t["achieved"] == pytest.approx(2.0 * d * 6.0e5 / 2e-3 / 1e12)
    assert t["frac"] == t["executed_frac"] == pytest.approx(t["achieved"] / 1500.0)
    assert t["required_tflops"] == pytest.approx(2.0 * d * 1.0e6 / 2e-3 / 1e12)
    assert t["computed_over_required"] == pytest.approx(0.6)
    rb = bench.kernel_rooflines(dict(stats, kernels={"refine_block": (1.0, 1)}), peaks)["refine_block"]
    assert rb["algorithmic_bytes_per_launch"] == 3000 * 16 + n * (w * 6 + 20) + 2500 * 8  # every sparse row once
    # no fused launches in this step: the tiled kernels are credited with every IVF row
    assert r["kmeans_tc_sparse"]["algorithmic_bytes_per_launch"] == 800 * (w * 6 + 6)
    stats["kernels"]["kmeans_fused"] = (1.0, 2)
    assert "kmeans_tc_sparse" not in bench.kernel_rooflines(stats, peaks)


def test_clock_sampler_without_nvidia_smi(bench, monkeypatch):
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first_sample(0.1)
    s.mark()
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"]
