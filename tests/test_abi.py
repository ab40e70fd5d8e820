"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/*.h declares, the
ctypes table covers them all, and the host mirror rejects CPU tensors (no fallback)."""
import os
import re

import pytest
import torch

from conftest import ROOT
import hd_yolo_b200 as hdy
from hd_yolo_b200 import _lib


def _declared():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"HDY_API\s+[\w\s\*]+?\b(hdy_\w+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 15
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.hdy_version()
    assert isinstance(lib.hdy_last_error(), bytes)


def test_argument_errors_are_detected_on_host():
    lib = _lib.load()
    # nl = 0 is invalid; nothing is launched, so this is safe without a GPU
    rc = lib.hdy_decode_concat(None, 0, 1, 3, 9, 0, None, None)
    assert rc == -1 and len(lib.hdy_last_error()) > 0
    rc = lib.hdy_nms_tiles(None, None, None, None, 1, 16, 0.5, 0.0, 0, 10, None, None, None, None, None, None, 0.0, None,
                           None, 0, None)
    assert rc == -1


def test_workspace_query():
    lib = _lib.load()
    assert lib.hdy_nms_workspace_bytes(8, 4096) == 0          # fits shared memory
    assert lib.hdy_nms_workspace_bytes(8, 25200) >= 8 * 32768 * 33


def test_cpu_tensors_are_rejected():
    spec = hdy.HeadSpec([[10, 13, 16, 30, 33, 23]] * 3, [8, 16, 32], nc=4)
    with pytest.raises(hdy.HdyError):
        hdy.nms_per_image(torch.zeros(1, 10, 9), nc=4)
    with pytest.raises(hdy.HdyError):
        hdy.non_max_suppression(torch.zeros(1, 10, 9))
    with pytest.raises(hdy.HdyError):
        hdy.compute_proposals([torch.zeros(1, 3, 4, 4, 9)] * 3, spec)
    with pytest.raises(hdy.HdyError):
        hdy.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)


def test_threshold_asserts_match_reference():
    # utils_general.py:312-313 / :442-443 raise AssertionError before touching the data
    with pytest.raises(AssertionError):
        hdy.nms_per_image(torch.zeros(1, 10, 9), nc=4, conf_thres=1.5)
    with pytest.raises(AssertionError):
        hdy.non_max_suppression(torch.zeros(1, 10, 9), iou_thres=-0.1)


def test_headspec_anchor_round_trip():
    from oracle import port
    from hd_yolo_b200 import synth
    spec = hdy.HeadSpec(synth.ANCHORS_4, synth.STRIDES_4, nc=7)
    _, ag = port.anchor_grids(synth.ANCHORS_4, synth.STRIDES_4)
    assert (torch.from_numpy(spec.anchor_grid) == ag).all()
    assert spec.na == 3 and spec.nl == 4 and spec.no == 12


def test_iou_threshold_rounding_modes():
    from hd_yolo_b200 import ops
    import numpy as np
    ops.set_iou_compare("cpu")
    assert ops._iou_thr_f32(0.3) < 0.3 <= float(np.float32(0.3))      # float32(0.3) > 0.3 -> rounded down
    assert ops._iou_thr_f32(0.45) == float(np.float32(0.45))          # float32(0.45) < 0.45 -> unchanged
    ops.set_iou_compare("cuda")
    assert ops._iou_thr_f32(0.3) == float(np.float32(0.3))
    ops.set_iou_compare("cpu")


def test_next_rows_reject_cpu_tensors_and_bad_arguments():
    with pytest.raises(hdy.HdyError):
        hdy.multiscale_roi_align([torch.zeros(1, 4, 8, 8)], torch.zeros(1, 5), None, [8])
    with pytest.raises(hdy.HdyError):
        hdy.box_iou(torch.zeros(2, 4), torch.zeros(3, 4))
    with pytest.raises(hdy.HdyError):
        hdy.match_predictions({'boxes': torch.zeros(1, 4), 'scores': torch.zeros(1), 'labels': torch.zeros(1)},
                              {'boxes': torch.zeros(1, 4), 'labels': torch.zeros(1)})
    lib = _lib.load()
    lv = (_lib.FeatureLevel * 1)()
    # sampling_ratio <= 0 (torchvision's adaptive grid) is not on the reference's path: rejected before any launch
    assert lib.hdy_multiscale_roi_align(lv, 1, 1, 4, None, None, 1, 14, 0, 0, None, None) == -1
    assert b"sampling_ratio" in lib.hdy_last_error()
    assert lib.hdy_multiscale_roi_align(lv, 1, 1, 4, None, None, 1, 17, 2, 0, None, None) == -1
    assert lib.hdy_match_pairs(None, None, None, 1, 70000, None, None, 70000, 0.5, 16, None, None, None, None,
                               None) == -1
