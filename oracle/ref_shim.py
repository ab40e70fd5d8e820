"""ORACLE (test infrastructure): import the real reference from /root/reference.

Works only in the build container (the reference does not travel to the GPU box); used by
oracle/make_golden.py to generate tests/golden/*.npz and by tests/test_oracle_vs_reference.py
(skipped when /root/reference is absent).  Recipe from SURVEY.md section 8c: a synthetic
``metayolo`` parent package so the reference's __init__ (matplotlib, skimage) never runs, and a stub
``torch_scatter`` (training-only import).
"""
import importlib.machinery
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("HD_YOLO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "metayolo", "models"))


def load():
    """Returns a namespace with the reference's hot-path symbols."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if "metayolo" not in sys.modules:
        pkg = types.ModuleType("metayolo")
        pkg.__path__ = [os.path.join(REF_ROOT, "metayolo")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("metayolo", None, is_package=True)
        pkg.LOGGER = logging.getLogger("yolov5")
        from packaging.version import parse

        pkg.check_version = lambda current="0", minimum="0", **kw: parse(current.split("+")[0]) >= parse(minimum)
        pkg.load_cfg = lambda cfg: cfg
        sys.modules["metayolo"] = pkg
        ts = types.ModuleType("torch_scatter")
        ts.scatter_max = None
        sys.modules["torch_scatter"] = ts
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        try:  # metayolo/models/metrics.py:11 imports pyplot at module level (plots only)
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    import warnings

    warnings.filterwarnings("ignore", message="torch.meshgrid")
    from metayolo.models import utils_general as ug
    from metayolo.models.yolo import Ensemble
    from metayolo.models.yolo_head import Detect
    import hnet.utils as hu
    from metayolo.models.metrics import APMeter
    from torchvision.models.detection.roi_heads import paste_masks_in_image

    ns = types.SimpleNamespace(
        nms_per_image=ug.nms_per_image, non_max_suppression=ug.non_max_suppression, xywh2xyxy=ug.xywh2xyxy,
        box_iou=ug.box_iou, scale_coords=ug.scale_coords, Detect=Detect, Ensemble=Ensemble,
        sliding_window_scanner=hu.sliding_window_scanner, split_by_sizes=hu.split_by_sizes,
        paste_masks_in_image=paste_masks_in_image, APMeter=APMeter,
    )
    return ns
