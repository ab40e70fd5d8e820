"""RoIAlign in front of the reference's mask head: host mirror of
``Detect.multiscale_roi_align`` (metayolo/models/yolo_head.py:279-299) and of the
``torchvision.ops.roi_align`` call it makes per level, over ``hdy_multiscale_roi_align``.

One launch covers every level (the level id routes each RoI to its feature map); the reference
loops over levels with ``torch.where(levels == i)`` and scatters into a zero tensor.  fp32, the
CPU op's operation order (bit-identical on finite inputs); CUDA tensors only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from . import _lib
from ._lib import HdyError, ptr
from .ops import DetectBatch, _call, _need_cuda, _stream

__all__ = ["multiscale_roi_align", "roi_align", "batch_rois", "compute_outputs"]


def _levels(features: Sequence[torch.Tensor], scales: Sequence[float], allow_channels_last: bool = False):
    """(level table, bs, channels, channels_last).  channels_last: every level is a [bs, C, h, w] tensor laid out
    [bs][h][w][C] in memory (torch.channels_last); only the tf32x3 mode reads that layout."""
    if len(features) != len(scales):
        raise HdyError("one spatial scale per feature level")
    if not 1 <= len(features) <= _lib.HDY_MAX_LEVELS:
        raise HdyError(f"1..{_lib.HDY_MAX_LEVELS} feature levels")
    arr = (_lib.FeatureLevel * len(features))()
    bs, ch = None, None
    plain = all(f.dim() == 4 and f.is_contiguous() for f in features)
    cl = (not plain) and allow_channels_last and all(
        f.dim() == 4 and f.is_contiguous(memory_format=torch.channels_last) for f in features)
    for i, f in enumerate(features):
        _need_cuda(f, f"features[{i}]")
        if not (plain or cl):
            raise HdyError(f"features[{i}] must be a contiguous [bs, C, h, w] tensor" +
                           (" (or all levels torch.channels_last)" if allow_channels_last else
                            "; channels_last is read by mode='tf32x3' only"))
        if bs is None:
            bs, ch = int(f.shape[0]), int(f.shape[1])
        elif (bs, ch) != (int(f.shape[0]), int(f.shape[1])):
            raise HdyError("feature levels disagree on batch size / channels")
        arr[i].data = f.data_ptr()
        arr[i].h, arr[i].w = int(f.shape[2]), int(f.shape[3])
        arr[i].spatial_scale = float(scales[i])
    return arr, bs, ch, cl


def multiscale_roi_align(features: List[torch.Tensor], boxes: torch.Tensor, levels: Optional[torch.Tensor],
                         strides: Sequence[float], output_size: int = 14, sampling_ratio: int = 2,
                         aligned: bool = False, mode: str = "exact") -> torch.Tensor:
    """Detect.multiscale_roi_align (yolo_head.py:279-299).

    mode "exact" (default): torchvision's operation order, bit-identical to its CPU op.  mode "tf32x3": the per-RoI
    [M*M x 36] x [36 x C] product on the tensor cores (tcgen05, 3xTF32 split, fp32 accumulation): ~1e-6 relative to
    the window's magnitude, several times faster; C must be a multiple of 64.  RoIs with tap windows over 6 x 6
    feature pixels are computed by the exact kernel in either mode.  In tf32x3 mode the levels may be
    torch.channels_last tensors (same shape, [bs][h][w][C] in memory): the window rows are then contiguous and the
    TMA loads several times cheaper -- the layout to keep the mask features in on B200.

    features: one [bs, C, h_i, w_i] tensor per level; boxes [K, 5] = (image index, x1, y1, x2, y2); levels [K] float
    level ids (the 'extra'[:, 0] column of nms_per_image); strides: buffer.stride per level (spatial_scale =
    1/stride, :295).  Returns [K, C, M, M] with M = output_size (the reference passes mask_output_size // 2); rows
    whose level id matches no level stay zero."""
    scales = [1.0 / float(s) for s in strides]
    arr, bs, ch, cl = _levels(features, scales, allow_channels_last=(mode == "tf32x3"))
    _need_cuda(boxes, "boxes")
    if boxes.dim() != 2 or boxes.shape[1] != 5:
        raise HdyError("boxes must be [K, 5] (image index, x1, y1, x2, y2)")
    K = int(boxes.shape[0])
    M = int(output_size)
    out = torch.empty((K, ch, M, M), dtype=torch.float32, device=boxes.device)
    if K == 0:
        return out
    boxes = boxes.contiguous()
    lv = None
    if levels is not None:
        _need_cuda(levels, "levels")
        if levels.shape != (K,):
            raise HdyError("levels must be [K]")
        lv = levels.contiguous()
    elif len(features) > 1:
        raise HdyError("levels is required with more than one feature level")
    if mode == "tf32x3":
        if ch % 64:
            raise HdyError(f"mode 'tf32x3' needs a multiple of 64 channels, got {ch}")
        fallback = torch.empty((K + 1,), dtype=torch.int32, device=boxes.device)
        _call("hdy_multiscale_roi_align_tf32x3", arr, len(features), bs, ch, int(cl), ptr(boxes), ptr(lv), K, M,
              int(sampling_ratio), int(bool(aligned)), ptr(out), ptr(fallback), _stream(), launches=2)
        return out
    if mode != "exact":
        raise HdyError(f"unknown roi_align mode {mode!r} ('exact' or 'tf32x3')")
    _call("hdy_multiscale_roi_align", arr, len(features), bs, ch, ptr(boxes), ptr(lv), K, M, int(sampling_ratio),
          int(bool(aligned)), ptr(out), _stream())
    return out


def roi_align(input: torch.Tensor, boxes: Union[torch.Tensor, List[torch.Tensor]], output_size, spatial_scale: float = 1.0,
              sampling_ratio: int = 2, aligned: bool = False, mode: str = "exact") -> torch.Tensor:
    """torchvision.ops.roi_align as the reference calls it (yolo_head.py:243, :294): Tensor[K, 5] boxes or a list of
    per-image Tensor[k_i, 4]; square output; sampling_ratio >= 1."""
    if isinstance(output_size, (tuple, list)):
        if len(output_size) != 2 or output_size[0] != output_size[1]:
            raise HdyError("only square outputs are on the reference's path")
        output_size = output_size[0]
    if isinstance(boxes, (list, tuple)):
        boxes = torch.cat([torch.nn.functional.pad(b, [1, 0], value=float(i)) for i, b in enumerate(boxes)])
    return multiscale_roi_align([input], boxes, None, [1.0 / float(spatial_scale)], int(output_size), sampling_ratio,
                                aligned, mode=mode)


def batch_rois(batch: DetectBatch, counts_host: Optional[Sequence[int]] = None):
    """The `proposals` / `levels` pair Detect.compute_outputs builds (yolo_head.py:320-328) from a DetectBatch:
    ([K, 5] image-index-padded boxes, [K] level ids), detections of image 0 first.  One D2H read of the counts unless
    they are given."""
    if counts_host is None:
        counts_host = batch.counts.cpu().tolist()
    rows, lv = [], []
    for i, k in enumerate(counts_host):
        if k:
            rows.append(torch.nn.functional.pad(batch.boxes[i, :k], [1, 0], value=float(i)))
            lv.append(batch.levels[i, :k])
    dev = batch.boxes.device
    if not rows:
        return torch.zeros((0, 5), dtype=torch.float32, device=dev), torch.zeros((0,), dtype=torch.float32, device=dev)
    return torch.cat(rows), torch.cat(lv)


def compute_outputs(dets: List[torch.Tensor], spec, features: Optional[List[torch.Tensor]] = None,
                    compute_masks: bool = True, seg_h=None, mask_indices: Optional[torch.Tensor] = None,
                    nms_params: Optional[dict] = None, mask_output_size: int = 28, aligned: bool = False,
                    multi_label: bool = False, hier_ops=None, layout: int = 0):
    """Detect.compute_outputs (yolo_head.py:301-355), composed: level concat + nms_per_image (:311-318), the
    `proposals` / `levels` pair (:320-328), multiscale_roi_align (:329), the mask head (:330, `seg_h`: the reference's
    own PyTorch module -- convolutions are not part of this package), sigmoid + per-label channel select (:331, 346-353),
    hierarchical scores + score / label select (:335-345).

    dets: the head's RAW level tensors ([bs,na,ny,nx,no] logits -- the decode of compute_proposals is fused in);
    features: the `seg` feature maps, one [bs, C, h_i, w_i] per level (needed with compute_masks); mask_indices:
    Detect.mask_indices ([1+nc] int64: channel of the mask head per label, < 0 = no mask); nms_params as
    Detect.get_nms_params returns them.  Returns the reference's List[Dict{'boxes','scores','labels'[, 'masks' [k,1,M,M]]}].
    The reference's `r['labels'].clamp(min=0.)` (:348) raises on torch >= 2 (a float clamp of an int64 tensor cannot
    index); the integer clamp it means is what runs here."""
    from .masks import mask_select
    from .ops import detect_postprocess
    p = {'conf_thres': 0.15, 'iou_thres': 0.45, 'max_det': 300}
    p.update(nms_params or {})
    batch = detect_postprocess(dets, spec, p['conf_thres'], p['iou_thres'], int(p['max_det']), layout=layout,
                               hier_ops=hier_ops)
    counts = batch.counts.cpu().tolist()       # the reference synchronises here too (n_obj_per_image, :318)
    results = batch.to_list(multi_label=multi_label, conf_thres=p['conf_thres'])
    if not compute_masks or multi_label or sum(counts) == 0:
        return results
    if features is None or seg_h is None or mask_indices is None:
        raise HdyError("compute_masks=True needs the seg feature maps, the seg_h module and mask_indices")
    proposals, levels = batch_rois(batch, counts)
    mask_features = multiscale_roi_align(features, proposals, levels, spec.strides, mask_output_size // 2, 2, aligned)
    mask_logits = seg_h(mask_features)                                   # the reference's mask head, PyTorch
    labels = torch.cat([r['labels'] for r in results])
    masks = mask_select(mask_logits.float().contiguous(), labels, mask_indices)   # sigmoid + channel pick + zeroing
    for r, m in zip(results, masks.split(counts, 0)):
        if len(r['boxes']):
            r['masks'] = m
    return results
