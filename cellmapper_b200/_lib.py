"""ctypes binding of libcellmapper_b200.so (the C ABI in include/cellmapper_b200.h).

There is no CPU fallback: if the shared library is missing or the device is not a B200 the
functions here raise ``RuntimeError`` (same convention as the reference's optional back-ends,
``src/cellmapper/check.py:44``).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("CM_LIBPATH") or os.path.join(HERE, "lib", "libcellmapper_b200.so")  # CM_LIBPATH: development builds (tools/)

# enum mirrors
F32, F64 = 0, 1
KERNELS = {"gaussian": 0, "scarches": 1, "inverse_distance": 2, "equal": 3}
DIST_SQRT_F64, DIST_SKLEARN_F32, DIST_SQUARED = 0, 1, 2
KNN_AUTO, KNN_EXACT_F64, KNN_TENSOR_EXHAUSTIVE = 0, 1, 2
EDGE_STATS_WORKSPACE_BYTES = 32768
KNN_ASSIGN_WORKSPACE_BYTES = 262144
SPGEMM_MAX_COLS = 40960
MMA_MAX_D, MMA_MAX_K = 128, 64  # limits of the tensor-core search (csrc/knn_internal.cuh)
SELECT_WORKSPACE_BYTES = 16384
MOMENTS = 8

_P = c_void_p
#: name -> (restype, argtypes); must list every function include/cellmapper_b200.h declares
SIGNATURES = {
    "cm_abi_version": (c_int, []),
    "cm_last_error": (c_char_p, []),
    "cm_device_check": (c_int, [c_int]),
    "cm_knn_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "cm_knn_search": (
        c_int,
        [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P, _P],
    ),
    "cm_knn_assign_reference": (
        c_int,
        [_P, c_int64, c_int64, c_int, c_int, c_int, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P],
    ),
    "cm_knn_search_cells": (
        c_int,
        [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P, _P, _P, _P],
    ),
    "cm_knn_merge_topk": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P, _P]),
    "cm_edge_stats": (c_int, [_P, _P, c_int64, _P, _P, _P, c_size_t, _P]),
    "cm_edge_kernel_to_csr": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P]),
    "cm_map_rows_fused": (
        c_int,
        [_P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, c_int64, c_int, c_int, _P, c_int64, _P],
    ),
    "cm_csr_row_normalize": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "cm_csr_col_sums": (c_int, [_P, _P, _P, c_int64, _P, _P]),
    "cm_vote_argmax": (c_int, [_P, _P, _P, c_int64, _P, c_int, _P, _P, _P, _P]),
    "cm_spmm_csr_dense": (c_int, [_P, _P, _P, c_int64, _P, c_int64, c_int, c_int, _P, c_int64, _P]),
    "cm_spgemm_count": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int32, _P, _P]),
    "cm_spgemm_fill": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, c_int, _P, c_int32, _P, _P, _P, _P]),
    "cm_spgemm_partition": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "cm_presence_workspace_bytes": (c_size_t, [c_int64, c_int, c_int64]),
    "cm_presence_scores": (c_int, [_P, _P, c_int64, c_int, _P, c_int64, c_int64, _P, c_int, _P, _P, _P, c_size_t, _P]),
    "cm_select_ranks": (c_int, [_P, c_int, c_int64, c_int64, _P, c_int, _P, _P, c_size_t, _P]),
    "cm_log1p_inplace": (c_int, [_P, c_int, c_int64, c_int64, _P]),
    "cm_clip_minmax_inplace": (c_int, [_P, c_int, c_int64, c_int64, c_double, c_double, c_double, c_double, c_int, _P]),
    "cm_expr_gene_sums": (
        c_int,
        [c_int, _P, _P, _P, c_int, c_int64, c_int64, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, _P, _P, _P],
    ),
    "cm_reverse_lists_workspace_bytes": (c_size_t, [c_int64]),
    "cm_reverse_lists": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, _P, c_size_t, _P]),
    "cm_jaccard_count": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P, _P, _P, _P, _P]),
    "cm_jaccard_fill": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P, _P, _P, c_int, _P, _P, _P, _P]),
    "cm_launch_count": (c_int64, []),
    "cm_profile_enable": (c_int, [c_int]),
    "cm_profile_last_knn_ms": (c_int, [_P]),
    "cm_debug_mma_tile": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P]),
}

#: development-build-only entry points (nvcc -DCM_DEV_PROBES); bound when present, never required
DEV_SIGNATURES = {
    "cm_debug_probe_flags": (c_int, [c_int]),
    "cm_debug_probe_prof": (c_int, [_P]),
}

_lib = None


def load(build_if_missing: bool = False):
    """Load the shared library (optionally building it first). Raises RuntimeError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        if build_if_missing:
            from . import build as _build

            _build.build()
        else:
            raise RuntimeError(
                f"{LIBPATH} not found. Build it with `python -m cellmapper_b200.build` "
                "(needs nvcc with sm_100a support). The 'b200' method has no CPU fallback."
            )
    lib = ctypes.CDLL(LIBPATH)
    for name, (restype, argtypes) in {**SIGNATURES, **DEV_SIGNATURES}.items():
        if not hasattr(lib, name):
            continue  # reported by tests/test_cabi.py; calling it raises AttributeError
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cm_abi_version() != 1:
        raise RuntimeError(f"ABI version mismatch: library reports {lib.cm_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().cm_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == 1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed (status {rc}): {msg}")


_checked_devices: set[int] = set()


def require_device(device_index: int) -> None:
    """Fail loudly unless CUDA + an sm_100 device + the native library are all present.
    The answer for a device cannot change within a process, so a passed check is remembered
    (cudaGetDeviceProperties costs ~2.5 ms per call)."""
    if int(device_index) in _checked_devices:
        return
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "method='b200' needs a CUDA device (NVIDIA B200, sm_100a); none is visible and there is no CPU fallback."
        )
    check(load().cm_device_check(int(device_index)), "cm_device_check")
    _checked_devices.add(int(device_index))
