"""ctypes binding of libmsvit.so (the C ABI declared in include/msvit.h).

There is no CPU fallback and no alternative backend: if the library is missing or a call
fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIB_PATH = os.path.join(CSRC_DIR, "libmsvit.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "msvit.h"))

F32, BF16 = 0, 1
DIST = {"rbf": 0, "cosine": 1, "normprod": 2}
DISC = {"kmeans": 0, "axis_align": 1}
MAX_EIG_BLOCK = 32

_lock = threading.Lock()
_lib = None

_c_int, _c_i64, _c_f32, _c_ptr = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

_SIGNATURES = {
    "msvit_version": (_c_int, []),
    "msvit_error_string": (ctypes.c_char_p, [_c_int]),
    "msvit_affinity_degree": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_f32,
                                       _c_f32, _c_ptr, _c_ptr, _c_ptr]),
    "msvit_ncut_eig": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _c_f32, _c_f32, _c_int, _c_ptr, _c_ptr, _c_ptr]),
    "msvit_ncut_fused": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_int,
                                  _c_int, _c_f32, _c_f32, _c_int, _c_int, _c_f32, _c_f32, _c_int, _c_ptr]),
    "msvit_ritz_kmeans": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int,
                                   _c_int, _c_int, _c_int, _c_int, _c_int, _c_f32, _c_int, _c_int, _c_ptr]),
    "msvit_discretise": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_int,
                                  _c_int, _c_f32, _c_int, _c_int, _c_ptr, _c_ptr]),
    "msvit_kmeans": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_int,
                              _c_int, _c_f32, _c_int, _c_ptr, _c_ptr]),
    "msvit_pool": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_build_segments": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_gather_rows": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_i64, _c_int, _c_ptr]),
    "msvit_compose_labels": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_gkm_assign": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr]),
    "msvit_gkm_workspace_bytes": (ctypes.c_size_t, [_c_i64, _c_int]),
    "msvit_gkm_sort": (_c_int, [_c_ptr, _c_i64, _c_int, _c_ptr, _c_ptr, _c_ptr, ctypes.c_size_t, _c_ptr]),
    "msvit_gkm_accumulate_workspace_bytes": (ctypes.c_size_t, [_c_int, _c_int]),
    "msvit_gkm_accumulate": (_c_int, [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_int, _c_int, _c_ptr,
                                      ctypes.c_size_t, _c_ptr]),
    "msvit_gkm_finalize": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_attention_mask": (_c_int, [_c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_cluster_key_sums": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr]),
    "msvit_cluster_attention_stats": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def build(verbose: bool = False) -> str:
    """Compile libmsvit.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libmsvit.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(there is no CPU or library fallback for the msvit kernels)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().msvit_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")
