"""Whole-slide post-processing pipeline: the composition the reference spreads over
``sliding_window_scanner`` (hnet/utils.py:37-62) -> per-tile ``Detect`` post-processing (metayolo/models/
yolo_head.py:160-181, 301-355) -> ``Detect.merge_outputs`` (:450-463) -> ``Ensemble.merge`` (metayolo/models/
yolo.py:165-204), kept on the device from head logits to slide-level verdicts and sharded over ranks by tile rows.

The backbone/neck/head convolutions are not part of this package: the caller hands over the head's raw level tensors
for each batch of tiles (``provider``).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import dist as hdist
from .ops import HeadSpec, detect_postprocess
from .slide import SlideAccumulator, _kept_in_order, merge_nms, sliding_window_scanner

__all__ = ["SlidePostprocessor"]


class SlidePostprocessor:
    """Post-processes this rank's share of a slide.

    provider(first_tile, last_tile) -> List[Tensor]: the head's level tensors ([bs,na,ny,nx,no], layout 0) for the
    global tiles [first_tile, last_tile) -- always a sub-range of ``self.tile_range``.
    """

    def __init__(self, spec: HeadSpec, image_size, roi_size, overlap: int, conf_thres: float, iou_thres: float,
                 max_det: int, cap: Optional[int] = None, batch: int = 128, rank: int = 0, world: int = 1,
                 group=None, device=None, capacity: Optional[int] = None):
        self.spec, self.conf, self.iou, self.max_det, self.cap = spec, conf_thres, iou_thres, max_det, cap
        self.batch, self.rank, self.world, self.group = batch, rank, world, group
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.rois = sliding_window_scanner(image_size, roi_size, overlap)            # [n_tiles, 4] host
        self.tile_range = hdist.shard_tile_rows(self.rois, world)[rank]
        t0, t1 = self.tile_range
        self.rois_dev = self.rois[t0:t1].to(self.device).contiguous()
        md = min(max_det, cap) if cap else max_det
        self.capacity = int(capacity) if capacity is not None else max((t1 - t0) * md, 1)
        self.acc = SlideAccumulator(self.capacity, self.device)

    def detect(self, provider: Callable[[int, int], List[torch.Tensor]]) -> None:
        """Per-tile post-processing of every own tile, appended in slide coordinates.  No host synchronisation."""
        t0, t1 = self.tile_range
        self.acc.reset()
        for a in range(t0, t1, self.batch):
            b = min(a + self.batch, t1)
            out = detect_postprocess(provider(a, b), self.spec, self.conf, self.iou, self.max_det, cap=self.cap)
            self.acc.append(out, self.rois_dev[a - t0:b - t0])

    def merge(self, ordered: bool = True) -> Dict[str, torch.Tensor]:
        """Slide-level Ensemble.merge over all ranks' detections.  Returns this rank's part: 'state' (verdict per own
        row), 'n' own rows, 'base' (first global row), and with ordered=True the own survivors in score-descending
        order ('boxes', 'scores', 'labels', 'index' = global row)."""
        n = self.acc.count()
        boxes, scores = self.acc.boxes[:n], self.acc.scores[:n]
        info: Dict[str, object] = {}
        if self.world > 1:
            res = hdist.merge_sharded(boxes, scores, self.conf, self.iou, group=self.group)
            state, base = res['state'], res['base']
            info = {'exchanges': res['exchanges'], 'seam_rows': res['seam_rows']}
        else:
            state, base = merge_nms(boxes, scores, self.conf, self.iou), 0
        out: Dict[str, object] = {'state': state, 'n': n, 'base': base, **info}
        if ordered:
            idx, ob, os_, ol = _kept_in_order(state, boxes, scores, self.acc.labels[:n], n)
            out.update({'boxes': ob, 'scores': os_, 'labels': ol, 'index': idx + base})
        return out

    def run(self, provider, ordered: bool = True) -> Dict[str, torch.Tensor]:
        self.detect(provider)
        return self.merge(ordered)
