"""Generate the golden fixtures in this directory by RUNNING the reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Writes
  get_dim.json       reference ``get_dim`` (spectrum.py:172-199) on a grid of settings
  binning.npz        reference ``_to_vector`` (spectrum.py:250-296) on seeded random peaks
  postprocess.npz    reference ``_linkage`` / ``_postprocess_cluster`` /
                     ``_get_cluster_group_idx`` (cluster.py:334-509) on seeded cases
  murmur.json        sklearn ``murmurhash3_32`` (the hash the published pipeline
                     called) on a seeded key set + sklearn's own known answers
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refexec  # noqa: E402


def main():
    assert refexec.available(), "needs /root/reference"
    rng = np.random.Generator(np.random.PCG64(20261017))

    # ---- get_dim
    spec = refexec.spectrum_functions()
    grid = [
        (101.0, 1500.0, 0.05), (101.0, 500.0, 0.05), (101.0, 1500.0, 0.02),
        (101.0, 1500.0, 0.1), (50.0, 2000.0, 0.05), (0.0, 1000.0, 1.0005079),
        (101.0, 1500.0, 0.005), (200.5, 1800.25, 0.03),
    ]
    rows = []
    for lo, hi, b in grid:
        n, s, e = spec["get_dim"](lo, hi, b)
        rows.append({"min_mz": lo, "max_mz": hi, "bin_size": b,
                     "vec_len": int(n), "start": float(s), "end": float(e)})
    with open(os.path.join(HERE, "get_dim.json"), "w") as fh:
        json.dump(rows, fh, indent=1)

    # ---- binning
    import numba as nb
    n_spec = 400
    counts = rng.integers(1, 51, n_spec)
    min_mz = rows[0]["start"]
    mzs, intens = nb.typed.List(), nb.typed.List()
    for c in counts:
        m = np.sort(rng.uniform(101.0, 1500.0, c)).astype(np.float32)
        # add values sitting exactly on/near bin edges
        if c > 3:
            k = np.float64(rng.integers(1, 27000))
            m[0] = np.float32(min_mz + 0.05 * k)
            m[1] = np.nextafter(np.float32(min_mz + 0.05 * k), np.float32(0))
            m = np.sort(m)
        mzs.append(m)
        intens.append(rng.random(c).astype(np.float32))
    data, indices, indptr = spec["_to_vector"](mzs, intens, min_mz, 0.05)
    np.savez_compressed(
        os.path.join(HERE, "binning.npz"),
        mz=np.concatenate(list(mzs)), intensity=np.concatenate(list(intens)),
        indptr=indptr.astype(np.int64), bins=indices, data=data,
        min_mz=np.float64(min_mz), bin_size=np.float64(0.05),
    )

    # ---- murmur
    from sklearn.utils import murmurhash3_32
    keys = np.r_[np.arange(0, 64), rng.integers(0, 2**31 - 1, 192)].astype(np.int32)
    mur = {
        "keys": keys.tolist(),
        "seed0": murmurhash3_32(keys, 0, True).tolist(),
        "seed42": murmurhash3_32(keys, 42, True).tolist(),
        # sklearn/utils/tests/test_murmurhash.py:11-23
        "kat": {"3,0": 847579505, "3,42": -1823081949, "3,0,positive": 847579505,
                "3,42,positive": 2471885347},
    }
    with open(os.path.join(HERE, "murmur.json"), "w") as fh:
        json.dump(mur, fh)

    # ---- postprocess
    cl = refexec.cluster_functions()
    # Environment shim (not an algorithm change): the reference's objmode block
    # declares ``int32[:]`` for ``np.unique(..., return_inverse=True)[1]``
    # (cluster.py:419-429), which current numpy returns as int64 and numba then
    # refuses to unbox.  Cast the inverse to int32 while the reference runs.
    _unique = np.unique

    def _unique32(*a, **k):
        out = _unique(*a, **k)
        if k.get("return_inverse") and isinstance(out, tuple):
            out = tuple(o.astype(np.int32) if i == 1 else o for i, o in enumerate(out))
        return out

    np.unique = _unique32
    link_cases, post_cases = [], []
    b3_vals = np.float32([500.000, 500.004, 500.009, 500.030, 500.031, 500.2])
    cases = [(b3_vals.astype(np.float64), "ppm", 20.0)]
    for _ in range(60):
        m = int(rng.integers(2, 40))
        centre = rng.uniform(300, 1500)
        mode = "ppm" if rng.random() < 0.6 else "Da"
        spread = centre * 20e-6 if mode == "ppm" else 0.05
        v = centre + rng.normal(0, spread * rng.uniform(0.2, 3.0), m)
        if rng.random() < 0.3:  # duplicates -> ties
            v[rng.integers(0, m, m // 3)] = v[0]
        if rng.random() < 0.3:  # float32-valued m/z as in the snapshot schema
            v = v.astype(np.float32).astype(np.float64)
        cases.append((v, mode, 20.0 if mode == "ppm" else 0.05))
    for v, mode, tol in cases:
        link = cl["_linkage"](v, mode)
        for rt_tol in (None, 5.0):
            labels = np.zeros(v.shape[0], np.int64)
            rts = rng.uniform(0, 30, v.shape[0])
            k = cl["_postprocess_cluster"](labels, v, rts, tol, mode, rt_tol, 2, 7)
            post_cases.append((v, rts, mode, tol, -1.0 if rt_tol is None else rt_tol, labels.copy(), int(k)))
        link_cases.append((v, mode, link))
    np.unique = _unique
    groups_in = np.array([-1, -1, 0, 0, 1, 2, 2, 2])
    groups_out = np.array(list(cl["_get_cluster_group_idx"](groups_in)))
    np.savez_compressed(
        os.path.join(HERE, "postprocess.npz"),
        n_link=len(link_cases), n_post=len(post_cases),
        groups_in=groups_in, groups_out=groups_out,
        **{f"link_v{i}": c[0] for i, c in enumerate(link_cases)},
        **{f"link_mode{i}": np.array(c[1]) for i, c in enumerate(link_cases)},
        **{f"link_out{i}": c[2] for i, c in enumerate(link_cases)},
        **{f"post_v{i}": c[0] for i, c in enumerate(post_cases)},
        **{f"post_rt{i}": c[1] for i, c in enumerate(post_cases)},
        **{f"post_mode{i}": np.array(c[2]) for i, c in enumerate(post_cases)},
        **{f"post_tol{i}": np.float64(c[3]) for i, c in enumerate(post_cases)},
        **{f"post_rttol{i}": np.float64(c[4]) for i, c in enumerate(post_cases)},
        **{f"post_labels{i}": c[5] for i, c in enumerate(post_cases)},
        **{f"post_k{i}": np.int64(c[6]) for i, c in enumerate(post_cases)},
    )
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
