"""Host-side mirror of the reference's box/NMS functions over the C-ABI kernels.

Every public function keeps the name, argument meaning, return structure and error
behaviour of the reference function it replaces (file:line relative to the
reference root are cited in each docstring).  PyTorch is used for device memory
and streams only; all arithmetic runs in ``libhdyolo_b200.so``.  CPU tensors are
rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import HdyError, Level, check, ptr

__all__ = [
    "HeadSpec",
    "set_iou_compare",
    "compute_proposals",
    "decode_concat",
    "nms_per_image",
    "non_max_suppression",
    "nms",
    "batched_nms",
    "detect_postprocess",
    "DetectBatch",
    "flatten_onehot_objects",
    "CapturedStep",
    "scratch_slot",
]

# ----------------------------------------------------------------------------- thresholds
_IOU_COMPARE = "cpu"


def set_iou_compare(mode: str) -> None:
    """Choose which torchvision build the IoU comparison mimics.

    torchvision's CPU kernel evaluates ``float(iou) > double(thr)``, its CUDA kernel
    ``float(iou) > float(thr)``.  They differ only when ``float(thr) > thr`` and an IoU lands
    exactly on ``float(thr)``.  "cpu" (default, matches the oracle) rounds the threshold down to
    fp32, "cuda" rounds to nearest; the kernels always compare in fp32.
    """
    global _IOU_COMPARE
    if mode not in ("cpu", "cuda"):
        raise ValueError("mode must be 'cpu' or 'cuda'")
    _IOU_COMPARE = mode


def _iou_thr_f32(thr: float) -> float:
    t = np.float32(thr)
    if _IOU_COMPARE == "cpu" and float(t) > float(thr):
        t = np.nextafter(t, np.float32(-np.inf), dtype=np.float32)
    return float(t)


def _conf_thr_f32(thr: float) -> float:
    # `tensor_fp32 > python_float` compares against float32(python_float) on every device
    return float(np.float32(thr))


# ----------------------------------------------------------------------------- plumbing
def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelProfile:
    """Optional per-call CUDA-event timing of the C-ABI launches (used by bench.py for the roofline
    numbers) and a launch counter.  Events are recorded on the stream the kernels are launched on."""

    # kernels launched per C-ABI call when it is not 1
    def __init__(self):
        self.enabled = False
        self.records = []   # (name, start_event, end_event)
        self.launches = 0

    def reset(self):
        self.records.clear()
        self.launches = 0

    def summary(self):
        """name -> (calls, total_ms); synchronises."""
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.records:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + s.elapsed_time(e))
        return out


profile = KernelProfile()


def _call(name: str, *args, launches: int = 1) -> None:
    fn = getattr(_lib.load(), name)
    if profile.enabled:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        profile.records.append((name, s, e))
    else:
        rc = fn(*args)
    profile.launches += launches
    check(rc, name)


class CapturedStep:
    """A fixed-shape post-processing step captured once into a CUDA graph and replayed.

    ``fn`` is any composition of this package's calls that does not synchronise with the host (candidate / mask
    capacities given, no ``to_list()``): its ~10 kernel launches and ~25 small allocations cost more host time than the
    kernels take on a B200 at 640-px batches, so a serving loop replays the graph instead.  ``fn`` must read its inputs
    from buffers that stay in place (write the next batch into them before calling); the returned object of the capture
    is handed back by every replay and is overwritten in place.

    The graph bakes in the addresses of the scratch buffers of its (slot, capture stream), which stay PINNED while the
    object lives: a call that would have to grow one of them raises instead of freeing memory the graph still writes
    to.  Scratch is keyed by stream, so eager work on other streams never touches them.  slot=None (default) takes a
    private slot nobody else uses."""

    _next_private = 1 << 20

    def __init__(self, fn, warmup: int = 3, slot: Optional[int] = None):
        if slot is None:
            slot = CapturedStep._next_private
            CapturedStep._next_private += 1
        self.slot = int(slot)
        cur = torch.cuda.current_stream()
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(cur)
        with scratch_slot(self.slot):
            with torch.cuda.stream(self.stream):
                for _ in range(max(warmup, 1)):      # grows every scratch buffer and sets kernel attributes
                    fn()
            cur.wait_stream(self.stream)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            before = profile.launches
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = fn()
        self.launches = profile.launches - before
        profile.launches = before
        self._pin = (self.slot, self.stream.cuda_stream)
        _pin_slot(self._pin, +1)

    def __del__(self):
        try:
            _pin_slot(self._pin, -1)
        except Exception:   # interpreter shutdown
            pass

    def __call__(self, stream: Optional[torch.cuda.Stream] = None):
        """Replay on the current stream (or `stream`).  Graphs captured with different `slot`s may be in flight at the
        same time on different streams; replays of one graph are ordered by the stream they are issued on."""
        if stream is None:
            self.graph.replay()
        else:
            with torch.cuda.stream(stream):
                self.graph.replay()
        profile.launches += self.launches
        return self.out


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HdyError(f"{name} must be a CUDA tensor: hd_yolo_b200 has no CPU fallback")
    if t.dtype != torch.float32:
        raise HdyError(f"{name} must be float32 (got {t.dtype}); the hot path is fp32 by contract")
    _same_device(t, name)


def _need_head_tensor(t: torch.Tensor, name: str) -> int:
    """Raw head outputs (logits, prototypes) may be fp32 or fp16 (a half() model, val_nuclei.py:115-116); returns the
    C-ABI dtype code.  fp16 values are widened on load: all arithmetic stays fp32."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise HdyError(f"{name} must be a CUDA tensor: hd_yolo_b200 has no CPU fallback")
    if t.dtype not in (torch.float32, torch.float16):
        raise HdyError(f"{name} must be float32 or float16 (got {t.dtype})")
    _same_device(t, name)
    return _lib.HDY_F16 if t.dtype == torch.float16 else _lib.HDY_F32


def _same_device(t: torch.Tensor, name: str) -> None:
    """Kernels are launched on the CURRENT device's current stream: a tensor of another GPU would be dereferenced
    there.  Enter ``torch.cuda.device(t.device)`` first."""
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise HdyError(f"{name} lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}: "
                       "wrap the call in `with torch.cuda.device(tensor.device):`")


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    return t if t.data_ptr() % 16 == 0 else t.clone()


_TLS = threading.local()          # the scratch slot is per thread (emulated ranks run on threads)
_PINNED: Dict[Tuple[int, int], int] = {}   # (slot, capture stream) -> live CUDA graphs that baked its buffers in
_PIN_LOCK = threading.Lock()


def _slot() -> int:
    return getattr(_TLS, "slot", 0)


def _pin_slot(slot, delta: int) -> None:
    with _PIN_LOCK:
        v = _PINNED.get(slot, 0) + delta
        if v > 0:
            _PINNED[slot] = v
        else:
            _PINNED.pop(slot, None)


class scratch_slot:
    """Context manager: work issued inside (by this thread) uses scratch buffer set `k`.  Steps that run concurrently
    (two CapturedStep graphs in flight, emulated ranks on threads) must use different slots; the default slot is 0."""

    def __init__(self, k: int):
        self.k = int(k)

    def __enter__(self):
        self.prev = _slot()
        _TLS.slot = self.k

    def __exit__(self, *exc):
        _TLS.slot = self.prev


class _Scratch:
    """Grow-only scratch buffers (candidate lists, NMS workspace), one set per (device, scratch slot, stream): work on
    two streams never shares a buffer by accident.  A (slot, stream) whose buffers a live CUDA graph has baked in
    (CapturedStep captures on a stream of its own) refuses to grow -- the old buffer would be freed under the graph."""

    def __init__(self):
        self._buf: Dict[Tuple[int, int, int, str], torch.Tensor] = {}

    def get(self, device: torch.device, name: str, nbytes: int) -> torch.Tensor:
        slot = _slot()
        dev = device.index if device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        key = (dev, slot, stream, name)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None and (slot, stream) in _PINNED:
                raise HdyError(f"scratch buffer '{name}' of slot {slot} would have to grow from {b.numel()} to {nbytes} "
                               "bytes, but a live CUDA graph (CapturedStep) has its address baked in: run this call "
                               "under another `scratch_slot`, or capture the graph for the larger shape")
            b = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b

    def clear(self):
        self._buf.clear()


_scratch = _Scratch()


@dataclass
class HeadSpec:
    """Static description of a Detect head (yolo_head.py:28-71): anchors in pixels per level,
    strides, class count and channels per row (no >= 5+nc; extra channels are carried as-is)."""

    anchors: Sequence[Sequence[float]]  # [nl][2*na] pixels, as in the model yaml
    strides: Sequence[float]
    nc: int
    no: Optional[int] = None

    def __post_init__(self):
        self.nl = len(self.strides)
        if len(self.anchors) != self.nl:
            raise ValueError("anchors and strides should have same length")
        self.na = len(self.anchors[0]) // 2
        if self.no is None:
            self.no = self.nc + 5
        if self.no < self.nc + 5:
            raise ValueError("no must be >= nc + 5")
        if self.na > _lib.HDY_MAX_ANCHORS or self.nl > _lib.HDY_MAX_LEVELS:
            raise ValueError("too many anchors / levels")
        # anchors are stored divided by stride and multiplied back at decode time
        # (yolo_head.py:59, :427) -- reproduce the same fp32 round trip.
        s = torch.tensor(list(self.strides), dtype=torch.float32)
        a = torch.tensor(self.anchors, dtype=torch.float32).view(self.nl, -1, 2) / s.view(-1, 1, 1)
        self.anchor_grid = (a * s.view(-1, 1, 1)).numpy()  # [nl, na, 2] pixels

    def rows_per_tile(self, shapes: Sequence[Tuple[int, int]]) -> int:
        return sum(self.na * ny * nx for ny, nx in shapes)

    def levels(self, dets: Sequence[torch.Tensor], layout: int = 0):
        if len(dets) != self.nl:
            raise ValueError(f"expected {self.nl} levels, got {len(dets)}")
        arr = (Level * self.nl)()
        bs = None
        shapes = []
        dtype = None
        for i, d in enumerate(dets):
            code = _need_head_tensor(d, f"dets[{i}]")
            if dtype is None:
                dtype = code
            elif dtype != code:
                raise HdyError("levels disagree on dtype")
            if code == _lib.HDY_F16 and layout != 0:
                raise HdyError("fp16 logits are read in layout 0 ([bs,na,ny,nx,no]) only")
            if not d.is_contiguous():
                raise HdyError(f"dets[{i}] must be contiguous")
            if layout == 0:
                if d.dim() != 5 or d.shape[1] != self.na or d.shape[4] != self.no:
                    raise HdyError(f"dets[{i}] must be [bs,{self.na},ny,nx,{self.no}], got {tuple(d.shape)}")
                b, _, ny, nx, _ = d.shape
            else:
                if d.dim() != 4 or d.shape[1] != self.na * self.no:
                    raise HdyError(f"dets[{i}] must be [bs,{self.na * self.no},ny,nx], got {tuple(d.shape)}")
                b, _, ny, nx = d.shape
            if bs is None:
                bs = b
            elif bs != b:
                raise HdyError("levels disagree on batch size")
            arr[i].logits = d.data_ptr()
            arr[i].dtype = code
            arr[i].ny, arr[i].nx = ny, nx
            arr[i].stride = float(self.strides[i])
            for a in range(self.na):
                arr[i].anchor_w[a] = float(self.anchor_grid[i, a, 0])
                arr[i].anchor_h[a] = float(self.anchor_grid[i, a, 1])
            shapes.append((ny, nx))
        return arr, bs, shapes


# ----------------------------------------------------------------------------- decode
def compute_proposals(dets: List[torch.Tensor], spec: HeadSpec) -> List[torch.Tensor]:
    """Detect.compute_proposals (metayolo/models/yolo_head.py:185-213): sigmoid + grid/anchor decode
    of every level, same shapes out ([bs,na,ny,nx,no])."""
    lib = _lib.load()
    levels, bs, _ = spec.levels(dets, 0)
    outs = [torch.empty_like(d, dtype=torch.float32) for d in dets]
    out_ptrs = (C.c_void_p * spec.nl)(*[o.data_ptr() for o in outs])
    _call("hdy_decode_levels", levels, spec.nl, bs, spec.na, spec.no, out_ptrs, _stream(), launches=spec.nl)
    return outs


def decode_concat(dets: List[torch.Tensor], spec: HeadSpec, layout: int = 0) -> torch.Tensor:
    """compute_proposals + the level-id pad and concat of Detect.compute_outputs
    (yolo_head.py:311-312): [bs, N, no+1]."""
    lib = _lib.load()
    levels, bs, shapes = spec.levels(dets, layout)
    N = spec.rows_per_tile(shapes)
    out = torch.empty((bs, N, spec.no + 1), dtype=torch.float32, device=dets[0].device)
    _call("hdy_decode_concat", levels, spec.nl, bs, spec.na, spec.no, layout, ptr(out), _stream(), launches=spec.nl)
    return out


# ----------------------------------------------------------------------------- NMS core
class _Cand:
    """Candidate lists of one batch (views into scratch)."""

    def __init__(self, device, bs: int, cap: int, with_cls: bool = False, tag: str = ""):
        self.bs, self.cap = bs, cap
        n = max(bs * cap, 1)
        self.keys = _scratch.get(device, tag + "keys", n * 8)
        self.boxes = _scratch.get(device, tag + "boxes", n * 16)
        self.cls = _scratch.get(device, tag + "cls", n * 4) if with_cls else None
        # counts[bs] followed by one status word
        self.counts = torch.empty(bs + 1, dtype=torch.int32, device=device)
        _call("hdy_zero_i32", ptr(self.counts), bs + 1, _stream())

    @property
    def status_ptr(self):
        return C.c_void_p(self.counts.data_ptr() + 4 * self.bs)


def _run_nms(cand: _Cand, iou_thres: float, max_det: int, class_offset: float = 0.0, max_nms: int = 0,
             want_cls: bool = False, gray_eps: float = 0.0, fragile_out: Optional[list] = None):
    lib = _lib.load()
    dev = cand.counts.device
    bs, cap = cand.bs, cand.cap
    md = max(1, min(int(max_det), cap))
    keep_idx = torch.empty((bs, md), dtype=torch.int32, device=dev)
    keep_slot = torch.empty((bs, md), dtype=torch.int32, device=dev)
    keep_box = torch.empty((bs, md, 4), dtype=torch.float32, device=dev)
    keep_score = torch.empty((bs, md), dtype=torch.float32, device=dev)
    keep_cls = torch.empty((bs, md), dtype=torch.float32, device=dev) if want_cls else None
    keep_counts = torch.empty(bs, dtype=torch.int32, device=dev)
    keep_frag = None
    if gray_eps > 0.0 and fragile_out is not None:
        keep_frag = torch.empty((bs, md), dtype=torch.uint8, device=dev)
        fragile_out.append(keep_frag)
    wbytes = lib.hdy_nms_workspace_bytes(bs, cap)
    ws = _scratch.get(dev, "nms_ws", wbytes) if wbytes else None
    _call("hdy_nms_tiles", ptr(cand.keys), ptr(cand.boxes), ptr(cand.cls), ptr(cand.counts), bs, cap,
                          _iou_thr_f32(iou_thres), float(class_offset), int(max_nms), md, ptr(keep_idx),
                          ptr(keep_slot), ptr(keep_box), ptr(keep_score), ptr(keep_cls), ptr(keep_counts),
                          float(gray_eps), ptr(keep_frag), ptr(ws), wbytes, _stream())
    return keep_idx, keep_slot, keep_box, keep_score, keep_cls, keep_counts, md


def _counts_to_host(cand: _Cand, keep_counts: torch.Tensor):
    """One D2H read: keep counts + the overflow status word (the reference syncs here too)."""
    both = torch.cat([keep_counts, cand.counts]).cpu()
    bs = cand.bs
    kc = both[:bs].tolist()
    status = int(both[2 * bs])
    if status & _lib.HDY_STATUS_OVERFLOW:
        need = int(both[bs:2 * bs].max())
        raise HdyError(f"candidate capacity overflow: a tile produced {need} candidates, cap={cand.cap}")
    return kc


def _check_thresholds(conf_thres, iou_thres):
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'


def nms_per_image(preds: torch.Tensor, nc: int, conf_thres: float = 0.25, iou_thres: float = 0.45,
                  max_det: int = 300) -> List[Dict[str, torch.Tensor]]:
    """nms_per_image (metayolo/models/utils_general.py:299-356).

    preds [bs, N, 5+nc+E] rows (cx, cy, w, h, obj, cls.., extra..).  Per image: xywh->xyxy, drop boxes
    with a side < 2 px, keep obj > conf_thres, class-agnostic greedy NMS on obj, first max_det.
    Returns [{'boxes': [k,4], 'scores': [k,1+nc], 'extra': [k,E]}] in NMS (score-descending) order.
    The reference's 10 s wall-clock early exit (:351-354) is intentionally not reproduced.
    """
    _check_thresholds(conf_thres, iou_thres)
    _need_cuda(preds, "preds")
    if preds.dim() != 3 or preds.shape[2] < 5 + nc:
        raise HdyError(f"preds must be [bs, N, >=5+nc], got {tuple(preds.shape)}")
    lib = _lib.load()
    preds = preds.contiguous()
    bs, N, row_len = preds.shape
    E = row_len - 5 - nc
    dev = preds.device
    if bs == 0:
        return []
    if N == 0:
        z = preds.new_zeros
        return [{'boxes': z((0, 4)), 'scores': z((0, 1 + nc)), 'extra': z((0, E))} for _ in range(bs)]
    cap = N
    cand = _Cand(dev, bs, cap)
    _call("hdy_filter_compact_preds", ptr(preds), bs, N, row_len, _conf_thr_f32(conf_thres), 2.0, cap,
                                     ptr(cand.keys), ptr(cand.boxes), ptr(cand.counts), cand.status_ptr,
                                     _stream())
    keep_idx, _, keep_box, _, _, keep_counts, md = _run_nms(cand, iou_thres, max_det)
    out_scores = torch.empty((bs, md, 1 + nc), dtype=torch.float32, device=dev)
    out_extra = torch.empty((bs, md, E), dtype=torch.float32, device=dev)
    _call("hdy_gather_preds", ptr(preds), bs, N, row_len, nc, ptr(keep_idx), ptr(keep_counts), md,
                             ptr(out_scores), ptr(out_extra) if E > 0 else None, _stream())
    kc = _counts_to_host(cand, keep_counts)
    return [{'boxes': keep_box[i, :k], 'scores': out_scores[i, :k], 'extra': out_extra[i, :k]}
            for i, k in enumerate(kc)]


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes=None, agnostic: bool = False, multi_label: bool = False, labels=(),
                        max_det: int = 300) -> List[torch.Tensor]:
    """non_max_suppression (metayolo/models/utils_general.py:423-523).

    prediction [bs, N, 5+nc].  obj > conf, conf = obj*cls, best class (or every class above conf when
    multi_label and nc > 1), optional class filter, at most 30000 boxes into NMS, class-offset
    (7680 px) greedy NMS unless agnostic, first max_det.  Returns a list of [k,6] = xyxy, conf, cls.
    merge-NMS is dead code in the reference (merge = False, :453) and is not provided; the 10 s
    early exit (:519-521) is not reproduced.
    """
    _check_thresholds(conf_thres, iou_thres)
    _need_cuda(prediction, "prediction")
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise HdyError(f"prediction must be [bs, N, 5+nc] with nc >= 1, got {tuple(prediction.shape)}")
    lib = _lib.load()
    dev = prediction.device
    bs = prediction.shape[0]
    nc = prediction.shape[2] - 5
    max_wh, max_nms = 7680, 30000
    multi_label = bool(multi_label) and nc > 1
    if bs == 0:
        return []
    if labels and any(len(lb) for lb in labels):
        # autolabelling: a-priori rows are appended after the confidence filter (:463-469); appending
        # them to the raw rows with obj = cls = 1 is equivalent (they pass every filter, keep their order)
        rows = max(len(lb) for lb in labels)
        extra = prediction.new_zeros((bs, rows, 5 + nc))
        for xi, lb in enumerate(labels):
            if len(lb):
                lb = lb.to(dev, torch.float32)
                extra[xi, :len(lb), :4] = lb[:, 1:5]
                extra[xi, :len(lb), 4] = 1.0
                extra[xi, torch.arange(len(lb), device=dev), lb[:, 0].long() + 5] = 1.0
        prediction = torch.cat([prediction, extra], 1)
    prediction = prediction.contiguous()
    N = prediction.shape[1]
    if N == 0:
        return [torch.zeros((0, 6), device=dev) for _ in range(bs)]
    cap = N * nc if multi_label else N
    cand = _Cand(dev, bs, cap, with_cls=True, tag="y")
    cmask = None
    if classes is not None:
        cm = torch.zeros(nc, dtype=torch.uint8)
        for c in classes:
            if 0 <= int(c) < nc and float(c) == int(c):
                cm[int(c)] = 1
        cmask = cm.to(dev)
    _call("hdy_filter_compact_yolo", ptr(prediction), bs, N, nc, _conf_thr_f32(conf_thres), int(multi_label),
                                    ptr(cmask), cap, ptr(cand.keys), ptr(cand.boxes), ptr(cand.cls),
                                    ptr(cand.counts), cand.status_ptr, _stream())
    _, _, keep_box, keep_score, keep_cls, keep_counts, md = _run_nms(
        cand, iou_thres, max_det, class_offset=0.0 if agnostic else float(max_wh), max_nms=max_nms, want_cls=True)
    out = torch.cat([keep_box, keep_score[..., None], keep_cls[..., None]], -1)
    kc = _counts_to_host(cand, keep_counts)
    return [out[i, :k] for i, k in enumerate(kc)]


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms (what the reference calls at utils_general.py:342, :507 and yolo.py:195):
    int64 indices of kept boxes, score-descending."""
    _need_cuda(boxes, "boxes")
    _need_cuda(scores, "scores")
    n = boxes.shape[0]
    if boxes.dim() != 2 or boxes.shape[1] != 4 or scores.shape != (n,):
        raise HdyError("boxes must be [n,4] and scores [n]")
    dev = boxes.device
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    lib = _lib.load()
    boxes = _aligned16(boxes.contiguous())
    scores = scores.contiguous()
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _call("hdy_make_keys", ptr(scores), 1, n, ptr(keys), _stream())
    cand = _Cand.__new__(_Cand)
    cand.bs, cand.cap = 1, n
    cand.keys, cand.boxes, cand.cls = keys, boxes, None
    cand.counts = torch.tensor([n, 0], dtype=torch.int32).to(dev)
    keep_idx, _, _, _, _, keep_counts, _ = _run_nms(cand, iou_threshold, n)
    k = int(keep_counts.item())
    return keep_idx[0, :k].to(torch.int64)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.batched_nms, coordinate-trick form (boxes + idxs*(max+1)), which is what
    torchvision uses on CUDA below 100 000 elements; used by the hnet heads (mask_rcnn.py:72, :192)."""
    _need_cuda(boxes, "boxes")
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    lib = _lib.load()
    dev = boxes.device
    boxes = _aligned16(boxes.contiguous())
    scores = scores.contiguous()
    off = float((boxes.max() + 1.0).item())  # max_coordinate + 1, fp32
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _call("hdy_make_keys", ptr(scores), 1, n, ptr(keys), _stream())
    cand = _Cand.__new__(_Cand)
    cand.bs, cand.cap = 1, n
    cand.keys, cand.boxes = keys, boxes
    cand.cls = idxs.to(torch.float32).contiguous()
    cand.counts = torch.tensor([n, 0], dtype=torch.int32).to(dev)
    keep_idx, _, _, _, _, keep_counts, _ = _run_nms(cand, iou_threshold, n, class_offset=off)
    k = int(keep_counts.item())
    return keep_idx[0, :k].to(torch.int64)


# ----------------------------------------------------------------------------- fused head path
@dataclass
class DetectBatch:
    """Struct-of-arrays result of the fused path; everything stays on the device.
    Row i of every array holds tile i's detections in slots [0, counts[i])."""

    boxes: torch.Tensor        # [bs, max_det, 4]  xyxy, tile coordinates
    scores_full: torch.Tensor  # [bs, max_det, 1+nc] after hierarchical_scores
    scores: torch.Tensor       # [bs, max_det]
    labels: torch.Tensor       # [bs, max_det] int64 (cls+1, or -100)
    levels: torch.Tensor       # [bs, max_det] float level id (the reference's 'extra' column)
    extra: Optional[torch.Tensor]  # [bs, max_det, no-5-nc] raw extra channels (mask coefficients)
    rows: torch.Tensor         # [bs, max_det] int32 source row inside the tile's [N] ordering
    counts: torch.Tensor       # [bs] int32
    cand_counts: torch.Tensor  # [bs+1] int32: candidates per tile after the filter, then status word
    max_det: int
    fragile: Optional[torch.Tensor] = None  # [bs, max_det] uint8: gray-zone flags (detect_postprocess(gray_eps=...))
    gray_eps: float = 0.0      # the perturbation the flags cover, and the IoU threshold they were produced for
    iou_thres: float = 0.0

    def to_list(self, multi_label: bool = False, conf_thres: float = 0.0) -> List[Dict[str, torch.Tensor]]:
        """The reference's List[Dict] (yolo_head.py:335-355); synchronises once."""
        both = torch.cat([self.counts, self.cand_counts]).cpu()
        bs = self.counts.shape[0]
        if int(both[-1]) & _lib.HDY_STATUS_OVERFLOW:
            raise HdyError(f"candidate capacity overflow (needed {int(both[bs:2 * bs].max())})")
        out = []
        for i, k in enumerate(both[:bs].tolist()):
            if multi_label:
                sc = self.scores_full[i, :k]
                out.append({'boxes': self.boxes[i, :k], 'scores': sc, 'labels': sc > conf_thres})
            else:
                out.append({'boxes': self.boxes[i, :k], 'scores': self.scores[i, :k], 'labels': self.labels[i, :k]})
        return out


def default_hier_ops(nc: int) -> List[Tuple[int, int]]:
    """Detect.build_hierarchical_tree's default {0: {1..nc}} -> cls_c *= obj (yolo_head.py:510-511)."""
    return [(c, 0) for c in range(1, nc + 1)]


def hier_ops_from_descendants(descendants: Dict[int, List[int]]) -> List[Tuple[int, int]]:
    """Flatten Detect.descendants (yolo_head.py:481-491) into ordered (dst, src) products,
    in the dict's iteration order exactly as hierarchical_scores (:473-479) applies them."""
    return [(v, k) for k, vs in descendants.items() for v in vs]


def detect_postprocess(dets: List[torch.Tensor], spec: HeadSpec, conf_thres: float = 0.15,
                       iou_thres: float = 0.45, max_det: int = 300, layout: int = 0,
                       cap: Optional[int] = None, hier_ops: Optional[List[Tuple[int, int]]] = None,
                       min_size: float = 2.0, gray_eps: float = 0.0) -> DetectBatch:
    """Fused Detect.compute_proposals + compute_outputs without masks
    (yolo_head.py:185-213, 301-345 -> utils_general.py:299-356): raw level logits in,
    final boxes / scores / labels out, no host synchronisation.

    cap bounds the candidates kept per tile after the confidence filter (default: all rows, which
    can never overflow); overflow is reported by DetectBatch.to_list().

    gray_eps > 0 (whole-slide pipeline): also report, per survivor, whether a neighbour's IoU could cross iou_thres
    once both boxes are shifted to slide coordinates and rounded by up to gray_eps per coordinate
    (DetectBatch.fragile; see hdy_nms_tiles in include/hd_yolo_b200.h).
    """
    _check_thresholds(conf_thres, iou_thres)
    lib = _lib.load()
    levels, bs, shapes = spec.levels(dets, layout)
    dev = dets[0].device
    N = spec.rows_per_tile(shapes)
    cap = N if cap is None else int(cap)
    nc, no = spec.nc, spec.no
    cand = _Cand(dev, bs, cap)
    cthr = _conf_thr_f32(conf_thres)
    _call("hdy_filter_compact_logits", levels, spec.nl, bs, spec.na, nc, no, layout, cthr, float(min_size), cap,
                                      ptr(cand.keys), ptr(cand.boxes), ptr(cand.counts), cand.status_ptr,
                                      _stream())
    frag: list = []
    keep_idx, _, keep_box, _, _, keep_counts, md = _run_nms(cand, iou_thres, max_det, gray_eps=gray_eps,
                                                            fragile_out=frag)
    scores_full = torch.empty((bs, md, 1 + nc), dtype=torch.float32, device=dev)
    lvl = torch.empty((bs, md), dtype=torch.float32, device=dev)
    ne = no - 5 - nc
    extra = torch.empty((bs, md, ne), dtype=torch.float32, device=dev) if ne > 0 else None
    score = torch.empty((bs, md), dtype=torch.float32, device=dev)
    label = torch.empty((bs, md), dtype=torch.int64, device=dev)
    ops = default_hier_ops(nc) if hier_ops is None else hier_ops
    flat = (C.c_int32 * (2 * len(ops)))(*[v for p in ops for v in p])
    if 1 + nc <= 32:
        _call("hdy_gather_select_logits", levels, spec.nl, bs, spec.na, nc, no, layout, ptr(keep_idx),
              ptr(keep_counts), md, flat, len(ops), cthr, ptr(scores_full), ptr(lvl), ptr(extra), ptr(score),
              ptr(label), _stream())
    else:
        _call("hdy_gather_logits", levels, spec.nl, bs, spec.na, nc, no, layout, ptr(keep_idx), ptr(keep_counts), md,
              ptr(scores_full), ptr(lvl), ptr(extra), _stream())
        _call("hdy_select_scores", ptr(scores_full), ptr(keep_counts), bs, md, nc, flat, len(ops), cthr, ptr(score),
              ptr(label), _stream())
    return DetectBatch(keep_box, scores_full, score, label, lvl, extra, keep_idx, keep_counts, cand.counts, md,
                       frag[0] if frag else None, float(gray_eps) if frag else 0.0, float(iou_thres))


def flatten_onehot_objects(x: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """flatten_onehot_objects (val_nuclei.py:34-48), caller-side glue for multi_label outputs: one row per set bit
    of the boolean 'labels' [k, 1+nc]; the row's label is the bit's column (column 0, objectness, becomes -100);
    boxes / masks are repeated, 'scores' [k, 1+nc] is read at the same (row, column).  Index arithmetic only, so it
    stays in PyTorch like the reference; no [k*(1+nc), ...] intermediate is materialised."""
    lab = x['labels']
    assert lab.dim() == 2, f"labels has shape: {lab.shape}, need an one hot tensor."
    row, col = torch.nonzero(lab > 0., as_tuple=True)       # row-major order == flatten()[keep] order
    res = {k: v for k, v in x.items()}
    labels = col.clone()
    labels[col == 0] = -100
    res['labels'] = labels
    res['boxes'] = x['boxes'][row]
    if 'scores' in x:
        res['scores'] = x['scores'][row, col]
    if 'masks' in x:
        res['masks'] = x['masks'][row]
    return res
