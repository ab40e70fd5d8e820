"""ORACLE (test infrastructure): ctypes wrapper over oracle/roi_align_core.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_roi.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.oracle_roi_align.restype = C.c_int
        _lib.oracle_roi_align.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_int64, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _lib


def roi_align(inp: np.ndarray, rois: np.ndarray, output_size: int, spatial_scale: float, sampling_ratio: int = 2,
              aligned: bool = False) -> np.ndarray:
    """torchvision.ops.roi_align(inp [N,C,H,W], rois [K,5], (M,M), spatial_scale, sampling_ratio, aligned)."""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    N, Cc, H, W = inp.shape
    K = rois.shape[0]
    out = np.zeros((K, Cc, output_size, output_size), dtype=np.float32)
    rc = _load().oracle_roi_align(inp.ctypes.data, N, Cc, H, W, rois.ctypes.data, K, float(np.float32(spatial_scale)),
                                  output_size, output_size, sampling_ratio, int(aligned), out.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle_roi_align")
    return out
