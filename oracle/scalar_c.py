"""ctypes access to oracle/_build/liboracle.so (TEST INFRASTRUCTURE): the plain-C
restatement of the paste recipe and raw moments, fast enough for full-size windows."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-C", HERE], check=True)
        L = C.CDLL(SO)
        L.oracle_paste_window.restype = C.c_int64
        L.oracle_paste_window.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_float, C.c_void_p]
        L.oracle_raw_moments.restype = None
        L.oracle_raw_moments.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def paste_window(mask: np.ndarray, box: np.ndarray, x_lo: int, x_hi: int, y_lo: int, y_hi: int,
                 thr: float = 0.5) -> np.ndarray:
    m = np.ascontiguousarray(mask, dtype=np.float32)
    b = np.ascontiguousarray(box, dtype=np.float32)
    out = np.zeros((y_hi - y_lo, x_hi - x_lo), dtype=np.uint8)
    lib().oracle_paste_window(m.ctypes.data, b.ctypes.data, x_lo, x_hi, y_lo, y_hi, thr,
                              out.ctypes.data)
    return out


def raw_moments(win: np.ndarray, x_off: int = 0, y_off: int = 0) -> np.ndarray:
    w = np.ascontiguousarray(win, dtype=np.uint8)
    m = np.zeros(10, dtype=np.int64)
    lib().oracle_raw_moments(w.ctypes.data, w.shape[1], w.shape[0], x_off, y_off, m.ctypes.data)
    return m
