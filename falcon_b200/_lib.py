"""ctypes binding of ``libfalcon_b200.so`` (the C ABI in ``include/falcon_b200.h``).

There is no CPU fallback: if the shared library is missing the import fails
loudly with the build command.  PyTorch is used by callers only to own device
buffers and streams; this module never touches torch.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfalcon_b200.so")

FLC_OK = 0
FLC_ERR_INVALID = -1
FLC_ERR_CUDA = -2
FLC_ERR_CAPACITY = -3
FLC_ERR_WORKSPACE = -4
FLC_ERR_UNSUPPORTED = -5
TOL_MODES = {"Da": 0, "ppm": 1}


SCALING = {None: 0, "root": 1, "log": 2, "rank": 3}  # flc_preprocess scaling codes


class CapacityError(RuntimeError):
    """A caller-provided device buffer was too small (FLC_ERR_CAPACITY)."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m falcon_b200.build` (needs nvcc). falcon_b200 has no CPU fallback."
        )
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_i32, _i64, _u32, _u64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
_f32, _f64, _sz = C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/falcon_b200.h one to one.
SIGNATURES = {
    "flc_last_error": (C.c_char_p, []),
    "flc_version": (C.c_int, []),
    "flc_launch_count": (_u64, []),
    "flc_reset_launch_count": (None, []),
    "flc_check_device": (C.c_int, [C.c_int]),
    "flc_profile_enable": (None, [C.c_int]),
    "flc_profile_reset": (None, []),
    "flc_profile_count": (C.c_int, []),
    "flc_profile_get": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(_f64), C.POINTER(C.c_int)]),
    "flc_get_dim": (C.c_int, [_f32, _f32, _f32, C.POINTER(_u32), C.POINTER(_f32), C.POINTER(_f32)]),
    "flc_hash_table": (C.c_int, [_u32, _u32, _u32, _p, _p]),
    "flc_vectorize": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f64, _f64, _u32, _u32, _u32, C.c_int,
                                _p, _i64, _p, _i64, _p, _p, _p, _p, _i32, _p, _p]),
    "flc_bucket_sort_workspace_bytes": (_sz, [_i64]),
    "flc_bucket_sort": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, C.POINTER(_i64), _i64, _p, _p, _sz, _p]),
    "flc_gather": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p]),
    "flc_scatter32": (C.c_int, [_p, _p, _i64, _p, _p]),
    "flc_ivf_plan": (C.c_int, [_p, _i64, _i32, C.c_int, _p, _p, _p, C.POINTER(_i64), C.POINTER(_i32),
                               C.POINTER(_i64), _p]),
    "flc_kmeans_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64, _i32, _u32]),
    "flc_kmeans_needs_tiled": (C.c_int, [_i64, _i64, _i32, _u32]),
    "flc_kmeans_train": (C.c_int, [_p, _p, _p, _i32, _p, _i64, _i64, _u32, _p, _i64, _p, _p, _i64, _i64, C.c_int,
                                   _p, _p, _i32, _p, _p, _p, _sz, _p]),
    "flc_ivf_assign": (C.c_int, [_p, _i64, _i64, _u32, _p, _i64, _p, _p, _p, _p, _i32,
                                 _p, _p, _i32, _p, _p, _p]),
    "flc_debug_kmeans_timing": (C.c_int, [_p]),
    "flc_preprocess_workspace_bytes": (_sz, [_i64, _i64]),
    "flc_preprocess": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _i32, _f32, _f32, _f32, _f32, _f32, _i32, C.c_int,
                                 _p, _p, _p, _p, C.POINTER(_i64), _p, _sz, _p]),
    "flc_medoids_workspace_bytes": (_sz, [_i64]),
    "flc_medoids": (C.c_int, [_p, _p, _p, _i64, _p, _i64, _p, _p, _sz, _p]),
    "flc_scan_workspace_bytes": (_sz, [_i64, _i64]),
    "flc_scan_pairs": (C.c_int, [_p, _i64, _i64, _u32, _p, _i64, _p, _p, _i32, _p, _f32, C.c_int,
                                 _p, _u64, _p, _p, _sz, _p]),
    "flc_knn_csr_workspace_bytes": (_sz, [_i64, _u64]),
    "flc_knn_csr": (C.c_int, [_p, _p, _u64, _p, _i64, _p, _p, _i32, _i64, _u32, _p, _p, _p, _p, _i32,
                              _f64, C.c_int, _f64, _i32, _i32, _f32, _p, _p, _p, _u64, _p,
                              C.POINTER(_i64), _p, _sz, _p]),
    "flc_dbscan_workspace_bytes": (_sz, [_i64]),
    "flc_dbscan": (C.c_int, [_p, _p, _p, _i64, _f32, _i32, _p, C.POINTER(_i64), _p, _i32, C.POINTER(_i32), _p,
                             _p, _sz, _p]),
    "flc_scatter_labels_peers": (C.c_int, [_p, _p, _i64, _p, _i64, _p, C.c_int, _i64, _p]),
    "flc_relabel_gathered": (C.c_int, [_p, C.c_int, _i64, _p, _p, _p]),
    "flc_split_workspace_bytes": (_sz, [_i64, C.c_int]),
    "flc_split_clusters": (C.c_int, [_p, _p, _p, _i64, _f64, C.c_int, _f64, _i32, C.c_int, _p,
                                     C.POINTER(_i64), _p, _p, _sz, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.flc_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Raise the Python exception matching a C-ABI error code.  Invalid
    arguments raise ``ValueError`` like the reference does
    (/root/reference/falcon/ms_io/ms_io.py:27-38, falcon/cluster/cluster.py:661-662)."""
    if rc == FLC_OK:
        return
    msg = last_error()
    if rc == FLC_ERR_INVALID:
        raise ValueError(msg)
    if rc == FLC_ERR_CAPACITY:
        raise CapacityError(msg)
    if rc == FLC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"falcon_b200 error {rc}: {msg}")


def ptr(t) -> C.c_void_p:
    """Raw pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(lib.flc_launch_count())


def reset_launch_count() -> None:
    lib.flc_reset_launch_count()


def profile_enable(on: bool) -> None:
    lib.flc_profile_enable(1 if on else 0)


def profile_reset() -> None:
    lib.flc_profile_reset()


def profile_summary() -> dict:
    """{kernel name: (total device ms, launches)} since the last reset."""
    out = {}
    for i in range(lib.flc_profile_count()):
        name = C.create_string_buffer(64)
        ms, k = C.c_double(0), C.c_int(0)
        check(lib.flc_profile_get(i, name, 64, C.byref(ms), C.byref(k)))
        out[name.value.decode()] = (ms.value, k.value)
    return out
