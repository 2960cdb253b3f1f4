"""ctypes binding of libpt_b200.so (include/prompt_tts_b200.h).

There is no fallback: if the shared library is missing or a call returns non-zero this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpt_b200.so")


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int64 * 4), ("stride", C.c_int64 * 4),
                ("kmajor", C.c_int32), ("batched", C.c_int32)]


class Segment(C.Structure):
    _fields_ = [("a_idx", C.c_int32), ("b_idx", C.c_int32), ("a_k0", C.c_int32), ("b_k0", C.c_int32),
                ("a_mn_shift", C.c_int32), ("b_mn_shift", C.c_int32), ("nk", C.c_int32), ("nrep", C.c_int32),
                ("rep_is_batch", C.c_int32), ("rep_c2_0", C.c_int32)]


class Gemm(C.Structure):
    _fields_ = [("a", Operand * 2), ("b", Operand * 2), ("seg", Segment * 8), ("nseg", C.c_int32),
                ("M", C.c_int32), ("N", C.c_int32), ("nz2", C.c_int32), ("nz3", C.c_int32),
                ("splitk", C.c_int32), ("block_n", C.c_int32),
                ("out", C.c_void_p), ("out_dtype", C.c_int32),
                ("out_stride_m", C.c_int64), ("out_stride_z2", C.c_int64), ("out_stride_z3", C.c_int64),
                ("alpha", C.c_float), ("bias", C.c_void_p), ("bias_z2", C.c_void_p), ("residual", C.c_void_p),
                ("res_stride_m", C.c_int64), ("res_stride_z2", C.c_int64), ("res_stride_z3", C.c_int64)]


OUT_BF16, OUT_F32, OUT_F32_ATOMIC_ADD = 0, 1, 2

_lib = None


class PtError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raise (never fall back) when it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PtError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU / PyTorch fallback for the hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.pt_last_error.restype = C.c_char_p
        _lib.pt_version.restype = C.c_int
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise PtError(f"{what}: rc={rc}: {lib().pt_last_error().decode()}")


def call(name: str, *args) -> None:
    """Call `pt_<name>` with ctypes-converted args and raise on a non-zero return."""
    fn = getattr(lib(), "pt_" + name)
    check(fn(*args), "pt_" + name)
