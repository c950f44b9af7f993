"""ctypes binding of libtris_sm100.so (C ABI: include/tris_sm100.h).

There is deliberately NO fallback: if the shared library is missing, cannot be loaded, or the device is
not an sm_100 GPU, every product entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRIS_LIB_PATH") or os.path.join(_HERE, "lib", "libtris_sm100.so")   # override: A/B builds of the same ABI

OP_K2D, OP_MN2D, OP_CONV = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_QUICKGELU = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("stats", C.c_void_p),
        ("a_mode", C.c_int32), ("b_mode", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldb", C.c_int32), ("ldd", C.c_int32),
        ("img_n", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32),
        ("taps", C.c_int32), ("flip", C.c_int32), ("wgrad", C.c_int32), ("b_tap_stride", C.c_int32),
        ("block_n", C.c_int32), ("split_k", C.c_int32), ("act", C.c_int32), ("out_dtype", C.c_int32),
        ("atomic", C.c_int32), ("max_ctas", C.c_int32),
        ("d_pre", C.c_void_p), ("dact_src", C.c_void_p),
        ("batch", C.c_int32), ("scale", C.c_float),
        ("a_batch_stride", C.c_int64), ("b_batch_stride", C.c_int64), ("d_batch_stride", C.c_int64),
        ("stats_parts", C.c_int32), ("stats_mode", C.c_int32),
        ("stats_y", C.c_void_p), ("stats_mu", C.c_void_p), ("mask_sc", C.c_void_p), ("mask_sh", C.c_void_p),
        ("splitk_ws", C.c_void_p),
        ("defer_reduce", C.c_int32), ("split_used", C.c_int32),
        ("d_norm", C.c_void_p), ("in_gamma", C.c_void_p), ("in_beta", C.c_void_p), ("in_mean", C.c_void_p), ("in_invstd", C.c_void_p),
        ("in_add", C.c_void_p), ("in_eps", C.c_float), ("in_mix", C.c_float), ("in_relu", C.c_int32),
        ("conv_halo", C.c_int32),
        ("res_bits", C.c_void_p), ("tstamp", C.c_void_p), ("residual_f32", C.c_int32),
    ]


class ReduceItem(C.Structure):
    _fields_ = [("ws", C.c_void_p), ("d", C.c_void_p), ("rows", C.c_int32), ("w", C.c_int32), ("ldd", C.c_int32),
                ("split", C.c_int32), ("accumulate", C.c_int32), ("pad_", C.c_int32)]


class TrisLibError(RuntimeError):
    pass


_lib = None
launch_count = 0  # number of kernel launches issued through this binding (bench.py's gpu_launches)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TrisLibError(f"{LIB_PATH} not found: run `python -m tris_b200.build` (no CPU / eager fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.tris_last_error.restype = C.c_char_p
        for name, fn in _lib.__dict__.items():
            pass
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise TrisLibError(f"{what} failed ({rc}): {lib().tris_last_error().decode()}")


def require_device():
    if not torch.cuda.is_available():
        raise TrisLibError("tris_b200 needs a CUDA device (sm_100a); there is no CPU path")
    check(lib().tris_check_device(), "tris_check_device")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def call(name: str, *args, launches: int = 1):
    """Invoke ``int name(..., stream)`` and raise on a non-zero return."""
    global launch_count
    fn = getattr(lib(), name)
    rc = fn(*args, stream_ptr())
    check(rc, name)
    launch_count += launches


def gemm_raw(desc: GemmDesc, launches: int = 1):
    global launch_count
    rc = lib().tris_gemm(C.byref(desc), stream_ptr())
    check(rc, "tris_gemm")
    launch_count += launches
