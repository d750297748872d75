"""ctypes binding of the C-ABI in include/lpformer_b200.h.

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "liblpformer_b200.so")

MODE = {"cn": 0, "1-hop": 1, "all": 2}
EPI_NONE, EPI_RELU, EPI_SIGMOID = 0, 1, 2

_p, _i64, _i32, _f32, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_int

# name -> (restype, argtypes); mirrors include/lpformer_b200.h line by line
SIGNATURES = {
    "lpf_abi_version": (_int, []),
    "lpf_last_error": (C.c_char_p, []),
    "lpf_device_ok": (_int, []),
    "lpf_select_count": (_int, [_p, _i64, _p, _p, _p, _p, _p, _f32, _f32, _f32, _int, _p, _p]),
    "lpf_scan_scratch_bytes": (_i64, [_i64]),
    "lpf_scan_counts": (_int, [_p, _i64, _p, _p, _p]),
    "lpf_select_fill": (_int, [_p, _i64, _p, _p, _p, _p, _p, _f32, _f32, _f32, _int, _p, _p, _p, _p, _p, _p]),
    "lpf_rpe_hidden": (_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _i32, _p, _i64, _p]),
    "lpf_gemm": (_int, [_p, _i64, _p, _i64, _p, _f32, _p, _i64, _i64, _i32, _i32, _int, _p]),
    "lpf_layernorm_act": (_int, [_p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i32, _int, _p]),
    "lpf_gather_links": (_int, [_p, _i64, _p, _i64, _i32, _p, _i64, _p, _i64, _p]),
    "lpf_attend_fused": (_int, [_p, _i64, _p, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p, _i32, _i32, _int, _int,
                                _p, _i64, _p, _p]),
    "lpf_gcn_spmm": (_int, [_p, _p, _p, _i64, _i64, _p, _i64, _p, _i32, _p, _i64, _p]),
}

_lib = None


class LpfError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LpfError(
            f"{LIB_PATH} not found: build the sm_100a extension first (python -m lpformer_b200.build). "
            "lpformer_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point; raise LpfError with lpf_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise LpfError(f"{name} failed ({rc}): {lib.lpf_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    """The current torch CUDA stream as a void*."""
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LpfError("lpformer_b200 kernels need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
