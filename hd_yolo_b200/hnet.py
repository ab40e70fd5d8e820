"""hnet multi-level heads: host-side mirror of the post-processing the reference's hnet/detection/mask_rcnn.py
delegates to torchvision, over csrc/rcnn.cu + the per-tile NMS kernel.

* ``box_decode``                  -- BoxCoder.decode            (hnet/detection/mask_rcnn.py:67, :192)
* ``rpn_filter_proposals``        -- RegionProposalNetwork.filter_proposals (mask_rcnn.py:72; thresholds
                                     hnet/detection/utils_det.py:27-29)
* ``roi_postprocess_detections``  -- RoIHeads.postprocess_detections (mask_rcnn.py:192; utils_det.py:49-51)
* ``maskrcnn_inference``          -- torchvision roi_heads.maskrcnn_inference (mask_rcnn.py:248)
* ``cross_level_merge``           -- rescale_outputs + merge_outputs + Ensemble.merge across magnifications
                                     (metayolo/models/yolo_head.py:450-471, yolo.py:165-204; the hnet-side
                                     composition is a TODO in the reference, hnet/hnet_new.py:275)

torchvision.ops.batched_nms has two arithmetically different forms: the coordinate trick (boxes + class * (max + 1),
one NMS) up to a size limit (box coordinates: 4000 on CPU, 100000 on CUDA in torchvision 0.26; read from the
INSTALLED torchvision at import, ``BATCHED_NMS_LIMITS``) and the class-separated "vanilla" form above it.  The two
round differently, so ``mode`` selects which arithmetic runs:
  "torchvision-cuda" (default) / "torchvision-cpu": torchvision's own size rule per image for that device -- what the
      reference executes (it runs on CUDA; at hnet's sizes, 1000 pre-NMS boxes per level, that is the trick).  Needs
      the candidate counts on the host.
  "vanilla": one batched launch over class-separated candidate lists (the throughput form; bit-parity with
      torchvision only where torchvision itself takes the vanilla form), "trick": the coordinate trick always.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import HdyError, ptr
from .ops import _Cand, _call, _need_cuda, _run_nms, _stream, batched_nms
from .masks import mask_select
from .slide import ensemble_merge, merge_outputs, rescale_outputs

__all__ = ["box_decode", "rpn_filter_proposals", "roi_postprocess_detections", "maskrcnn_inference",
           "cross_level_merge"]

XFORM_CLIP = math.log(1000.0 / 16)
_MODES = ("vanilla", "trick", "torchvision-cpu", "torchvision-cuda")


def _torchvision_limits() -> Dict[str, int]:
    """(cpu, cuda) element-count limits of torchvision.ops.batched_nms' coordinate trick, from the installed source."""
    lim = {"cpu": 4000, "cuda": 100000}
    try:
        import inspect
        import re
        from torchvision.ops import boxes as _b
        m = re.search(r"numel\(\)\s*>\s*\(\s*([\d_]+)\s*if[^)]*?cpu[^)]*?else\s*([\d_]+)\s*\)",
                      inspect.getsource(_b.batched_nms))
        if m:
            lim = {"cpu": int(m.group(1).replace("_", "")), "cuda": int(m.group(2).replace("_", ""))}
    except Exception:   # no torchvision / source not available: the documented 0.26 values
        pass
    return lim


BATCHED_NMS_LIMITS = _torchvision_limits()


def _a16(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def box_decode(rel_codes: torch.Tensor, boxes, weights=(1.0, 1.0, 1.0, 1.0), bbox_xform_clip: float = XFORM_CLIP,
               boxes_rows: Optional[int] = None) -> torch.Tensor:
    """BoxCoder.decode (torchvision models/detection/_utils.py): rel_codes [R, C*4], boxes = tensor [R, 4] or a list
    of per-image tensors (concatenated, as torchvision does) -> [R, C, 4].  boxes_rows < R re-uses the boxes
    cyclically (RPN: the same anchors for every image)."""
    if isinstance(boxes, (list, tuple)):
        boxes = torch.cat(list(boxes), 0)
    _need_cuda(rel_codes, "rel_codes")
    _need_cuda(boxes, "boxes")
    R = rel_codes.shape[0]
    if rel_codes.dim() != 2 or rel_codes.shape[1] % 4:
        raise HdyError("rel_codes must be [R, C*4]")
    Cn = rel_codes.shape[1] // 4
    rows = boxes.shape[0] if boxes_rows is None else int(boxes_rows)
    if boxes_rows is None and rows != R:
        raise HdyError("boxes and rel_codes disagree on the number of rows")
    out = torch.empty((R, Cn, 4), dtype=torch.float32, device=rel_codes.device)
    if R:
        _call("hdy_rcnn_decode", ptr(_a16(rel_codes)), ptr(_a16(boxes)), R, Cn, rows, float(weights[0]),
              float(weights[1]), float(weights[2]), float(weights[3]), float(bbox_xform_clip), ptr(out), _stream())
    return out


def _img_wh(image_shapes, dev) -> torch.Tensor:
    return torch.tensor([[float(w), float(h)] for (h, w) in image_shapes], dtype=torch.float32).to(dev)


def _key_scores(keys: torch.Tensor) -> torch.Tensor:
    """fp32 score stored in the high word of candidate keys ((~orderable(score) << 32) | index, csrc/hdy_common.cuh)."""
    o = (~(keys >> 32)) & 0xffffffff                                    # orderable(score)
    bits = torch.where(o >= 0x80000000, o & 0x7fffffff, (~o) & 0xffffffff)
    return torch.where(bits >= 0x80000000, bits - (1 << 32), bits).to(torch.int32).view(torch.float32)


def _cut(cand: _Cand, max_det: int):
    """Score-ordered cut of per-image candidate lists: hdy_nms_tiles with iou 2 sorts and caps only."""
    return _run_nms(cand, 2.0, max_det, want_cls=True)


def _class_separated(cand: _Cand, n_img: int, group: int, iou: float, top_n: int):
    """NMS on (image, class) lists, survivors regrouped per image, score-ordered cut."""
    keep_idx, _, keep_box, keep_score, keep_cls, keep_counts, md = _run_nms(cand, iou, cand.cap, want_cls=True)
    img = _Cand(cand.counts.device, n_img, cand.cap * group, with_cls=True, tag="hnet_img")
    _call("hdy_regroup_kept", ptr(keep_idx), ptr(keep_box), ptr(keep_score), ptr(keep_cls), ptr(keep_counts),
          n_img * group, group, md, img.cap, ptr(img.keys), ptr(img.boxes), ptr(img.cls), ptr(img.counts),
          img.status_ptr, _stream())
    return _cut(img, top_n)


def _per_image_rule(cand: _Cand, n_img: int, iou: float, top_n: int, mode: str):
    """torchvision's own batched_nms per image (size rule: coordinate trick up to 4000 / 100000 box coordinates,
    class-separated above).  cand holds one list per image with classes."""
    limit = {"torchvision-cpu": BATCHED_NMS_LIMITS["cpu"], "torchvision-cuda": BATCHED_NMS_LIMITS["cuda"],
             "trick": 1 << 62, "vanilla": -1}[mode]
    counts = cand.counts[:n_img].cpu().tolist()
    dev = cand.counts.device
    keys = cand.keys.view(torch.int64)[:n_img * cand.cap].view(n_img, cand.cap)
    boxes = cand.boxes.view(torch.float32)[:n_img * cand.cap * 4].view(n_img, cand.cap, 4)
    cls = cand.cls.view(torch.float32)[:n_img * cand.cap].view(n_img, cand.cap)
    out = []
    for i, n in enumerate(counts):
        n = min(n, cand.cap)
        if n == 0:
            out.append((boxes.new_zeros((0, 4)), boxes.new_zeros((0,)), boxes.new_zeros((0,))))
            continue
        # restore torchvision's row order (the key's low word is the row index in its flattened arrays)
        order = torch.argsort(keys[i, :n] & 0xffffffff)
        b, c = boxes[i, :n][order].contiguous(), cls[i, :n][order].contiguous()
        s = _key_scores(keys[i, :n][order])
        if n * 4 > limit:        # class-separated: NMS per class, then scores[keep].sort(descending)
            kept = []
            from .ops import nms
            for cval in torch.unique(c).tolist():
                m = torch.nonzero(c == cval).flatten()
                kept.append(m[nms(b[m].contiguous(), s[m].contiguous(), iou)])
            k = torch.cat(kept)
            k = k[torch.argsort(s[k], descending=True, stable=True)]
        else:
            k = batched_nms(b, s, c, iou)
        k = k[:top_n]
        out.append((b[k], s[k], c[k]))
    return out


def rpn_filter_proposals(proposals: torch.Tensor, objectness: torch.Tensor, image_shapes: Sequence[Tuple[int, int]],
                         num_anchors_per_level: Sequence[int], pre_nms_top_n: int = 1000, post_nms_top_n: int = 1000,
                         nms_thresh: float = 0.7, score_thresh: float = 0.0, min_size: float = 1e-3,
                         mode: str = "torchvision-cuda") -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """RegionProposalNetwork.filter_proposals (torchvision models/detection/rpn.py): proposals [N, A, 4] (decoded
    anchors), objectness [N, A] (or [N*A, 1]) logits -> (boxes per image, scores per image), best first."""
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {_MODES}")
    _need_cuda(proposals, "proposals")
    _need_cuda(objectness, "objectness")
    N, A = proposals.shape[0], proposals.shape[1]
    dev = proposals.device
    objectness = objectness.reshape(N, -1).contiguous()
    if objectness.shape[1] != A or sum(num_anchors_per_level) != A:
        raise HdyError("objectness / num_anchors_per_level disagree with proposals")
    nl = len(num_anchors_per_level)
    sizes = (C.c_int32 * nl)(*[int(v) for v in num_anchors_per_level])
    keys = torch.empty((N * A,), dtype=torch.int64, device=dev)
    _call("hdy_rpn_level_keys", ptr(objectness), N, A, sizes, nl, ptr(keys), _stream())
    from .slide import sort_keys
    sort_keys(keys)
    per_level = mode == "vanilla"
    topk = [min(int(v), pre_nms_top_n) for v in num_anchors_per_level]
    cap = max(topk) if per_level else sum(topk)
    cand = _Cand(dev, N * nl if per_level else N, cap, with_cls=True, tag="rpn")
    _call("hdy_rpn_topk_compact", ptr(keys), ptr(_a16(proposals)), N, A, sizes, nl, int(pre_nms_top_n),
          ptr(_img_wh(image_shapes, dev)), float(min_size), float(score_thresh), int(per_level), cap, ptr(cand.keys),
          ptr(cand.boxes), ptr(cand.cls), ptr(cand.counts), cand.status_ptr, _stream())
    if per_level:
        _, _, kb, ks, _, kc, _ = _class_separated(cand, N, nl, nms_thresh, post_nms_top_n)
        kc = kc.cpu().tolist()
        return [kb[i, :k] for i, k in enumerate(kc)], [ks[i, :k] for i, k in enumerate(kc)]
    res = _per_image_rule(cand, N, nms_thresh, post_nms_top_n, mode)
    return [r[0] for r in res], [r[1] for r in res]


def roi_postprocess_detections(class_logits: torch.Tensor, box_regression: torch.Tensor,
                               proposals: List[torch.Tensor], image_shapes: Sequence[Tuple[int, int]],
                               box_weights=(10.0, 10.0, 5.0, 5.0), score_thresh: float = 0.05,
                               nms_thresh: float = 0.5, detections_per_img: int = 100, min_size: float = 1e-2,
                               mode: str = "torchvision-cuda"):
    """RoIHeads.postprocess_detections (torchvision models/detection/roi_heads.py): class_logits [R, C],
    box_regression [R, C*4], proposals = per-image [r_i, 4] -> (boxes, scores, labels) per image, best first."""
    if mode not in _MODES:
        raise ValueError(f"mode must be one of {_MODES}")
    _need_cuda(class_logits, "class_logits")
    dev = class_logits.device
    R, Cn = class_logits.shape
    n_img = len(proposals)
    pred_boxes = box_decode(box_regression, proposals, box_weights)                      # [R, C, 4]
    scores = torch.empty((R, Cn), dtype=torch.float32, device=dev)
    if R:
        _call("hdy_softmax_rows", ptr(class_logits.contiguous()), R, Cn, ptr(scores), _stream())
    rows = [int(p.shape[0]) for p in proposals]
    offs = torch.tensor([0] + rows, dtype=torch.int64).cumsum(0).to(torch.int32).to(dev)
    per_class = mode == "vanilla"
    group = Cn - 1
    cap = max(max(rows), 1) if per_class else max(max(rows) * group, 1)
    cand = _Cand(dev, n_img * group if per_class else n_img, cap, with_cls=True, tag="roi")
    if R and Cn > 1:
        _call("hdy_rcnn_filter_compact", ptr(pred_boxes), ptr(scores), ptr(offs), ptr(_img_wh(image_shapes, dev)),
              n_img, R, Cn, float(score_thresh), float(min_size), int(per_class), cap, ptr(cand.keys),
              ptr(cand.boxes), ptr(cand.cls), ptr(cand.counts), cand.status_ptr, _stream())
    if per_class:
        _, _, kb, ks, kcls, kc, _ = _class_separated(cand, n_img, group, nms_thresh, detections_per_img)
        kc = kc.cpu().tolist()
        return ([kb[i, :k] for i, k in enumerate(kc)], [ks[i, :k] for i, k in enumerate(kc)],
                [kcls[i, :k].to(torch.int64) for i, k in enumerate(kc)])
    res = _per_image_rule(cand, n_img, nms_thresh, detections_per_img, mode)
    return [r[0] for r in res], [r[1] for r in res], [r[2].to(torch.int64) for r in res]


def maskrcnn_inference(x: torch.Tensor, labels: List[torch.Tensor]) -> List[torch.Tensor]:
    """torchvision roi_heads.maskrcnn_inference: sigmoid of the mask logits [K, C, M, M], channel = label of each
    box -> per-image [k_i, 1, M, M]."""
    lab = torch.cat(labels)
    ident = torch.arange(x.shape[1], dtype=torch.int64, device=x.device)
    probs = mask_select(x, lab, ident)
    return list(probs.split([int(l.shape[0]) for l in labels], 0))


def cross_level_merge(levels: Sequence[Tuple[float, List[Dict[str, torch.Tensor]]]], nms_params: Dict[str, float],
                      task_id: str = "det") -> Dict[str, torch.Tensor]:
    """Hierarchical merge across magnifications: for every (scale, tiles) pair the tiles' detections are brought to
    slide coordinates (Detect.merge_outputs, yolo_head.py:450-463), rescaled to the common frame
    (rescale_outputs, :465-471), and everything is merged by Ensemble.merge (yolo.py:165-204)."""
    parts = []
    for scale, tiles in levels:
        m = merge_outputs(tiles)
        rescale_outputs(m, scale)
        parts.append({task_id: m})
    return ensemble_merge(parts, nms_params)[task_id]
