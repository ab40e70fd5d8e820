"""ORACLE (test infrastructure): ctypes wrapper over oracle/nms_core.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_nms.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_nms.restype = C.c_int64
        _lib.oracle_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    return _lib


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.empty(n, dtype=np.int64)
    k = _load().oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(iou_threshold), keep.ctypes.data)
    if k < 0:
        raise MemoryError("oracle_nms")
    return keep[:k]
