"""ctypes loader of the in-tree native library (libplslam.so, sm_100a).

There is no fallback: if the library is missing or cannot be loaded, importing the product API fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libplslam.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
KL_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                     ("response", "<f4"), ("size", "<f4"), ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"),
                     ("sx_oct", "<f4"), ("sy_oct", "<f4"), ("ex_oct", "<f4"), ("ey_oct", "<f4"), ("length", "<f4"),
                     ("num_pixels", "<i4")])
assert KP_DTYPE.itemsize == 28 and KL_DTYPE.itemsize == 68

PL_OK, PL_ERR_ARG, PL_ERR_EMPTY, PL_ERR_CUDA, PL_ERR_CAPACITY, PL_ERR_STATE = 0, -1, -2, -3, -4, -5


class PlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"plslam error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path.")
        _lib = C.CDLL(LIB_PATH)
        _lib.pl_last_error.restype = C.c_char_p
        _lib.pl_build_info.restype = C.c_char_p
        _lib.pl_orb_scale_factor.restype = C.c_float
        _lib.pl_orb_scale_factor.argtypes = [C.c_void_p]
        for name in ("pl_orb_stream", "pl_match_stream", "pl_line_stream"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_void_p
                getattr(_lib, name).argtypes = [C.c_void_p]
    return _lib


def check(rc: int):
    if rc != PL_OK:
        raise PlError(rc, lib().pl_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array or a raw integer device address."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)
