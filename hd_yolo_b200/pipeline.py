"""Whole-slide post-processing pipeline: the composition the reference spreads over
``sliding_window_scanner`` (hnet/utils.py:37-62) -> per-tile ``Detect`` post-processing (metayolo/models/
yolo_head.py:160-181, 301-355) -> ``Detect.merge_outputs`` (:450-463) -> ``Ensemble.merge`` (metayolo/models/
yolo.py:165-204), kept on the device from head logits to slide-level verdicts and sharded over ranks by tile rows.

The backbone/neck/head convolutions are not part of this package: the caller hands over the head's raw level tensors
for each batch of tiles (``provider``).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
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
                 group=None, device=None, capacity: Optional[int] = None, interior_shortcut: bool = True,
                 streams: int = 1):
        self.spec, self.conf, self.iou, self.max_det, self.cap = spec, conf_thres, iou_thres, max_det, cap
        self.batch, self.rank, self.world, self.group = batch, rank, world, group
        # streams > 1: consecutive tile batches are post-processed on alternating side streams (own scratch slot
        # each), so one batch's latency-bound per-tile NMS overlaps the next batch's HBM-bound filter; the appends stay
        # in tile order through an event chain
        self.n_streams = max(1, int(streams))
        self._side: List[torch.cuda.Stream] = []
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.rois = sliding_window_scanner(image_size, roi_size, overlap)            # [n_tiles, 4] host
        self.tile_range = hdist.shard_tile_rows(self.rois, world)[rank]
        t0, t1 = self.tile_range
        self.rois_dev = self.rois[t0:t1].to(self.device).contiguous()
        md = min(max_det, cap) if cap else max_det
        self.capacity = int(capacity) if capacity is not None else max((t1 - t0) * md, 1)
        self.acc = SlideAccumulator(self.capacity, self.device)
        # Interior shortcut of the slide merge: a detection strictly inside its tile's core can only be suppressed by
        # a survivor of its own tile, and only if rounding `box + tile origin` to fp32 lifts their IoU over the
        # threshold.  The per-tile NMS flags exactly those pairs (gray zone, half an ulp of the largest slide
        # coordinate per box coordinate), so everything else in the interior is KEPT without a pair test -- exactly.
        h, w = (image_size, image_size) if isinstance(image_size, (int, float)) else (image_size[0], image_size[1])
        top = float(max(h, w)) * 1.01 + 64.0
        self.gray_eps = float(np.spacing(np.float32(top))) / 2.0
        # the cell margins of the NMS binning cover a rounding of < 0.004 px (slides up to 131 072 px) -- beyond
        # that, and for degenerate thresholds, every row takes the unconditional path
        self.shortcut = bool(interior_shortcut) and self.gray_eps <= 0.004 and iou_thres >= 0.05

    def detect(self, provider: Callable[[int, int], List[torch.Tensor]]) -> None:
        """Per-tile post-processing of every own tile, appended in slide coordinates.  No host synchronisation."""
        t0, t1 = self.tile_range
        self.acc.reset()
        if self.n_streams == 1:
            for a in range(t0, t1, self.batch):
                b = min(a + self.batch, t1)
                out = detect_postprocess(provider(a, b), self.spec, self.conf, self.iou, self.max_det, cap=self.cap,
                                         gray_eps=self.gray_eps if self.shortcut else 0.0)
                self.acc.append(out, self.rois_dev[a - t0:b - t0])
            return
        from .ops import scratch_slot
        with torch.cuda.device(self.device):
            while len(self._side) < self.n_streams:
                self._side.append(torch.cuda.Stream())
            cur = torch.cuda.current_stream()
            for s in self._side:
                s.wait_stream(cur)
            prev = None
            for i, a in enumerate(range(t0, t1, self.batch)):
                b = min(a + self.batch, t1)
                s = self._side[i % self.n_streams]
                with torch.cuda.stream(s), scratch_slot(8 + i % self.n_streams):
                    dets = provider(a, b)
                    out = detect_postprocess(dets, self.spec, self.conf, self.iou, self.max_det, cap=self.cap,
                                             gray_eps=self.gray_eps if self.shortcut else 0.0)
                    if prev is not None:
                        s.wait_event(prev)           # appends in tile order: the accumulator's cursor is shared
                    self.acc.append(out, self.rois_dev[a - t0:b - t0])
                    prev = torch.cuda.Event()
                    prev.record(s)
            for s in self._side:
                cur.wait_stream(s)

    def merge(self, ordered: bool = True) -> Dict[str, torch.Tensor]:
        """Slide-level Ensemble.merge over all ranks' detections.  Returns this rank's part: 'state' (verdict per own
        row), 'n' own rows, 'base' (first global row), and with ordered=True the own survivors in score-descending
        order ('boxes', 'scores', 'labels', 'index' = global row)."""
        n = self.acc.count()
        boxes, scores = self.acc.boxes[:n], self.acc.scores[:n]
        info: Dict[str, object] = {}
        if self.world > 1:
            kw = {}
            if self.shortcut:
                # tile ids of the accumulator are local; the core table covers the whole slide
                t0 = self.tile_range[0]
                tl = self.acc.tile[:n]
                tg = torch.where(tl >= 0, tl + t0, ~((~tl) + t0))
                from .slide import dirty_tiles, tile_cores
                if getattr(self, "_cores_all", None) is None:
                    self._cores_all = tile_cores(self.rois).to(self.device)
                    self._rois_all = self.rois.to(self.device).contiguous()
                margin, far_boxes, far_tile, far_count = self.acc.overhang()
                # far-reaching boxes of every rank can touch any rank's tiles: gather the (short) lists
                import torch.distributed as dist
                nf = min(int(far_count.item()), int(far_boxes.shape[0]))
                overflow = int(far_count.item()) > int(far_boxes.shape[0])
                cnts = [torch.empty((2,), dtype=torch.int64, device=self.device) for _ in range(self.world)]
                dist.all_gather(cnts, torch.tensor([nf, int(overflow)], dtype=torch.int64, device=self.device),
                                group=self.group)
                sizes = [int(c[0]) for c in cnts]
                pay = torch.cat([far_boxes[:nf], (far_tile[:nf] + t0).to(torch.float32)[:, None]], 1)
                parts = hdist._pad_gather(pay, sizes, self.group)
                allf = torch.cat(parts) if sum(sizes) else torch.zeros((0, 5), dtype=torch.float32, device=self.device)
                cap = max(int(allf.shape[0]), 1)
                fb = torch.zeros((cap, 4), dtype=torch.float32, device=self.device)
                ft = torch.zeros((cap,), dtype=torch.int32, device=self.device)
                fb[:allf.shape[0]] = allf[:, :4]
                ft[:allf.shape[0]] = allf[:, 4].to(torch.int32)
                total = int(allf.shape[0]) + (cap + 1 if any(int(c[1]) for c in cnts) else 0)   # > cap: all dirty
                fc = torch.tensor([total], dtype=torch.int32, device=self.device)
                kw = dict(tile_id=tg.contiguous(), cores=self._cores_all, margin=margin,
                          dirty=dirty_tiles(fb, ft, fc, self._rois_all))
            res = hdist.merge_sharded(boxes, scores, self.conf, self.iou, group=self.group, **kw)
            state, base = res['state'], res['base']
            info = {'exchanges': res['exchanges'], 'seam_rows': res['seam_rows']}
        else:
            state, base = self.acc.verdicts(self.conf, self.iou, interior_shortcut=self.shortcut)[:n], 0
        out: Dict[str, object] = {'state': state, 'n': n, 'base': base, **info}
        if ordered:
            idx, ob, os_, ol = _kept_in_order(state, boxes, scores, self.acc.labels[:n], n)
            out.update({'boxes': ob, 'scores': os_, 'labels': ol, 'index': idx + base})
        return out

    def run(self, provider, ordered: bool = True) -> Dict[str, torch.Tensor]:
        self.detect(provider)
        return self.merge(ordered)
