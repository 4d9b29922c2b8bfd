"""Build ``libfalcon_b200.so`` in-tree with nvcc for sm_100a.

``python -m falcon_b200.build`` (or ``__graft_entry__.build()``).  The library
links the CUDA runtime statically and resolves the one driver entry point it
needs (``cuTensorMapEncodeTiled``) at run time, so it loads on a machine
without a GPU driver -- nvcc cross-compiles here, the GPU box only loads it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libfalcon_b200.so")
SOURCES = ["api.cu", "vectorize.cu", "bucket.cu", "scan.cu", "scan_tc.cu", "refine.cu", "kmeans.cu", "kmeans_tc.cu", "dbscan.cu", "medoids.cu", "preprocess.cu", "gather.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wno-deprecated-declarations",
    "-diag-suppress", "1444",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libfalcon_b200.so")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = True) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "falcon_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                print("[falcon_b200.build]", " ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append((src, out))
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"== {s} ==\n{o}" for s, o in failed))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", LIB, *objs,
               "-Xlinker", "--exclude-libs,ALL"]
        if verbose:
            print("[falcon_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
    print(LIB)
