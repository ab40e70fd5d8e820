"""Whole-slide post-processing pipeline: the composition the reference spreads over
``sliding_window_scanner`` (hnet/utils.py:37-62) -> per-tile ``Detect`` post-processing (metayolo/models/
yolo_head.py:160-181, 301-355) -> ``Detect.merge_outputs`` (:450-463) -> ``Ensemble.merge`` (metayolo/models/
yolo.py:165-204), kept on the device from head logits to slide-level verdicts (and masks of the kept detections) and
sharded over ranks by tile rows.

The backbone/neck/head convolutions are not part of this package: the caller hands over the head's raw level tensors
for each batch of tiles (``provider``) and, for masks, the prototype maps (``proto_provider``).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import dist as hdist
from ._lib import HdyError
from .masks import PackedMasks, SlideMaskBuilder
from .ops import DetectBatch, HeadSpec, detect_postprocess
from .slide import (SlideAccumulator, _gather_ordered, _kept_in_order, _order_keys, sliding_window_scanner,
                    tile_cores)

__all__ = ["SlidePostprocessor"]


class SlidePostprocessor:
    """Post-processes this rank's share of a slide.

    provider(first_tile, last_tile) -> List[Tensor]: the head's level tensors ([bs,na,ny,nx,no], layout 0; fp32 or
    fp16) for the global tiles [first_tile, last_tile) -- always a sub-range of ``self.tile_range``.
    proto_provider(first_tile, last_tile) -> Tensor [bs, nm, mh, mw] (fp32 or fp16): the prototype maps of the same
    tiles (north-star mask variant); needs a head with nm extra channels (``spec.no == 5 + nc + nm``).

    world > 1: ranks own bands of tile rows; collectives go over ``comm`` (default: torch.distributed on ``group``;
    ``hd_yolo_b200.dist.ThreadGroup`` emulates the ranks inside one process).
    """

    def __init__(self, spec: HeadSpec, image_size, roi_size, overlap: int, conf_thres: float, iou_thres: float,
                 max_det: int, cap: Optional[int] = None, batch: int = 128, rank: int = 0, world: int = 1,
                 group=None, device=None, capacity: Optional[int] = None, interior_shortcut: bool = True,
                 streams: int = 1, comm=None, seam_cap: int = 65536):
        self.spec, self.conf, self.iou, self.max_det, self.cap = spec, conf_thres, iou_thres, max_det, cap
        self.batch, self.rank, self.world, self.group = batch, rank, world, group
        # streams > 1: consecutive tile batches are post-processed on alternating side streams (own scratch slot
        # each), so one batch's latency-bound per-tile NMS overlaps the next batch's HBM-bound filter; the appends stay
        # in tile order through an event chain
        self.n_streams = max(1, int(streams))
        self._side: List[torch.cuda.Stream] = []
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.image_size = (image_size, image_size) if isinstance(image_size, (int, float)) else \
            (image_size[0], image_size[1])
        self.roi_size = (roi_size, roi_size) if isinstance(roi_size, (int, float)) else (roi_size[0], roi_size[1])
        self.rois = sliding_window_scanner(image_size, roi_size, overlap)            # [n_tiles, 4] host
        self.tile_range = hdist.shard_tile_rows(self.rois, world)[rank]
        t0, t1 = self.tile_range
        self.rois_dev = self.rois[t0:t1].to(self.device).contiguous()
        md = min(max_det, cap) if cap else max_det
        self.capacity = int(capacity) if capacity is not None else max((t1 - t0) * md, 1)
        self.comm = None
        self.backend = None
        room = 0
        if world > 1:
            self.comm = comm if comm is not None else hdist.TorchDistComm(group)
            if self.comm.rank != rank or self.comm.world != world:
                raise HdyError("rank / world disagree with the communicator")
            self.backend = hdist.DeviceSeamBackend(self.device, rank, world, conf_thres, iou_thres, seam_cap=seam_cap)
            room = 2 * self.backend.rep_cap          # the replicas live behind the own rows of the accumulator
        self.acc = SlideAccumulator(self.capacity + room, self.device)
        self.acc._top = float(self.rois.abs().max()) if self.rois.numel() else 0.0
        # Interior shortcut of the slide merge: a detection strictly inside its tile's core can only be suppressed by
        # a survivor of its own tile, and only if rounding `box + tile origin` to fp32 lifts their IoU over the
        # threshold.  The per-tile NMS flags exactly those pairs (gray zone, half an ulp of the largest slide
        # coordinate per box coordinate), so everything else in the interior is KEPT without a pair test -- exactly.
        h, w = self.image_size
        top = float(max(h, w)) * 1.01 + 64.0
        self.gray_eps = float(np.spacing(np.float32(top))) / 2.0
        # the cell margins of the NMS binning cover a rounding of < 0.004 px (slides up to 131 072 px) -- beyond
        # that, and for degenerate thresholds, every row takes the unconditional path
        self.shortcut = bool(interior_shortcut) and self.gray_eps <= 0.004 and iou_thres >= 0.05
        self._cores_all = self._rois_all = None
        self._batches: List[Tuple[int, int, DetectBatch, torch.Tensor]] = []   # (a, b, outputs, tile_offsets)
        self.keep_batches = False

    # ------------------------------------------------------------------------------------------ per-tile part
    def _post(self, dets):
        return detect_postprocess(dets, self.spec, self.conf, self.iou, self.max_det, cap=self.cap,
                                  gray_eps=self.gray_eps if self.shortcut else 0.0)

    def _append(self, out: DetectBatch, a: int, b: int) -> None:
        t0 = self.tile_range[0]
        self.acc.append(out, self.rois_dev[a - t0:b - t0], rois_host=self.rois[a:b])
        # a candidate list that outgrew `cap` truncated its tile: carried into the accumulator's status word, which
        # merge() reads (nobody calls to_list() on the batches of a slide)
        self.acc.status.bitwise_or_(out.cand_counts[-1:])
        if self.keep_batches:       # the mask pass needs the tile-local boxes and the coefficients again (only those)
            keep = DetectBatch(out.boxes, None, None, None, None, out.extra, None, out.counts, out.cand_counts,
                               out.max_det)
            self._batches.append((a, b, keep, self.acc.tile_offsets[-1]))

    def detect(self, provider: Callable[[int, int], List[torch.Tensor]], keep_batches: bool = False) -> None:
        """Per-tile post-processing of every own tile, appended in slide coordinates.  No host synchronisation.
        keep_batches: retain every batch's outputs for ``masks()``."""
        t0, t1 = self.tile_range
        self.acc.reset()
        self._batches = []
        self.keep_batches = bool(keep_batches)
        if self.n_streams == 1:
            for a in range(t0, t1, self.batch):
                b = min(a + self.batch, t1)
                self._append(self._post(provider(a, b)), a, b)
            return
        from .ops import scratch_slot, _slot
        base_slot = _slot()
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
                with torch.cuda.stream(s), scratch_slot(base_slot * 16 + 8 + i % self.n_streams):
                    out = self._post(provider(a, b))
                    if prev is not None:
                        s.wait_event(prev)           # appends in tile order: the accumulator's cursor is shared
                    self._append(out, a, b)
                    prev = torch.cuda.Event()
                    prev.record(s)
            for s in self._side:
                cur.wait_stream(s)

    # ------------------------------------------------------------------------------------------ slide-level part
    def merge(self, ordered: bool = True) -> Dict[str, torch.Tensor]:
        """Slide-level Ensemble.merge over all ranks' detections.  Returns this rank's part: 'state' (verdict per own
        row), 'n' own rows, 'base' (first global row), and with ordered=True the own survivors in score-descending
        order ('boxes', 'scores', 'labels', 'index' = global row)."""
        n = self.acc.count()                # the one read in front of the merge: rows appended (+ overflow check)
        acc = self.acc
        info: Dict[str, object] = {}
        if self.world > 1:
            be = self.backend
            kw, over = {}, None
            if self.shortcut:
                acc.check_shortcut(self.iou)
                if self._cores_all is None:     # tile ids of the accumulator are local; the tables cover the slide
                    self._cores_all = tile_cores(self.rois).to(self.device)
                    self._rois_all = self.rois.to(self.device).contiguous()
            hook, box = None, {}
            if ordered:
                cnt = be.meta[hdist.M_KEPT:hdist.M_KEPT + 1]

                def hook(state):            # enqueued in front of the single host read of the merge
                    box['keys'] = _order_keys(state, acc.scores, n, cnt)
            while True:
                acc.reserve(n + be.rep_cap)         # (no-op unless the seam blocks had to grow)
                if self.shortcut:
                    kw = dict(tile_id=acc.tile, tile_base=self.tile_range[0], cores=self._cores_all,
                              rois_all=self._rois_all)
                    over = acc.overhang() if acc.rois else None
                try:
                    res = hdist.seam_merge(self.comm, be, acc.boxes, acc.scores, n, over, after_finish=hook, **kw)
                    break
                except hdist.SeamOverflow as e:     # raised on every rank alike: grow the blocks and repeat
                    be.set_seam_cap(int(e.needed * 1.25) + 1024)
            state, base = res['state'].clone(), res['base']       # (the backend's buffer is reused by the next merge)
            info = {'exchanges': res['exchanges'], 'seam_rows': res['seam_rows']}
            out: Dict[str, object] = {'state': state, 'n': n, 'base': base, **info}
            if ordered:
                idx, ob, os_, ol = _gather_ordered(box['keys'], cnt, res['meta'][hdist.M_KEPT], n, acc.boxes,
                                                   acc.scores, acc.labels)
                out.update({'boxes': ob, 'scores': os_, 'labels': ol, 'index': idx + base})
            return out
        state, base = acc.verdicts(self.conf, self.iou, interior_shortcut=self.shortcut)[:n], 0
        out = {'state': state, 'n': n, 'base': base}
        if ordered:
            idx, ob, os_, ol = _kept_in_order(state, acc.boxes[:n], acc.scores[:n], acc.labels[:n], n)
            out.update({'boxes': ob, 'scores': os_, 'labels': ol, 'index': idx + base})
        return out

    def masks(self, proto_provider: Callable[[int, int], torch.Tensor], state: torch.Tensor,
              words_per_row: float = 56.0, upsample: bool = True) -> PackedMasks:
        """process_mask (bit-packed, upsampled to tile pixels) for the rows the slide-level merge KEPT: a second pass
        over the tiles, reading every prototype map once.  Needs ``detect(..., keep_batches=True)``.  Returns one
        PackedMasks over this rank's slide rows (windows in slide pixels); ``.check()`` reads the overflow flag."""
        if not self._batches and self.tile_range[1] > self.tile_range[0]:
            raise HdyError("masks() needs detect(..., keep_batches=True)")
        nm = self.spec.no - 5 - self.spec.nc
        if nm <= 0:
            raise HdyError("masks() needs a head with mask coefficients (spec.no > 5 + nc)")
        n = int(state.shape[0])
        # canvas of the windows: the scanner clips the last row / column of tile WINDOWS to the image, but a head still
        # sees a full tile there, so a mask may reach up to one tile beyond the clipped window's origin
        canvas = (int(self.rois[:, 1].max()) + int(self.roi_size[0]), int(self.rois[:, 0].max()) + int(self.roi_size[1]))
        bld = SlideMaskBuilder(n, int(n * words_per_row) + 4096, canvas, self.device)
        t0 = self.tile_range[0]
        if self.n_streams == 1:
            for a, b, out, offs in self._batches:
                bld.add_batch(proto_provider(a, b), out.extra, out.boxes, out.counts, offs,
                              self.rois_dev[a - t0:b - t0], state, self.roi_size, upsample=upsample)
            return bld.finish()
        # The region kernel waits on HBM, the upsample kernel on issue slots: batches alternate over the side streams,
        # so one batch's upsample overlaps the next batch's regions.  Windows / offsets are prepared on the calling
        # stream in batch order (the word cursor is shared); a scratch pair is re-used only after its kernels finished.
        from .ops import scratch_slot, _slot
        base_slot = _slot()
        with torch.cuda.device(self.device):
            while len(self._side) < self.n_streams:
                self._side.append(torch.cuda.Stream())
            cur = torch.cuda.current_stream()
            done: List[Optional[torch.cuda.Event]] = [None] * self.n_streams
            for i, (a, b, out, offs) in enumerate(self._batches):
                lane = i % self.n_streams
                side = self._side[lane]
                if done[lane] is not None:
                    cur.wait_event(done[lane])
                ready = torch.cuda.Event()
                with torch.cuda.stream(side), scratch_slot(base_slot * 16 + 8 + lane):
                    protos = proto_provider(a, b)            # (a host-resident provider copies on this stream)
                pshape = tuple(protos.shape[2:])
                prep = bld.prepare(out.boxes, out.counts, offs, self.rois_dev[a - t0:b - t0], state, pshape,
                                   self.roi_size, upsample, lane=lane)
                ready.record(cur)
                with torch.cuda.stream(side), scratch_slot(base_slot * 16 + 8 + lane):
                    side.wait_event(ready)
                    bld.run(prep, protos, out.extra, out.boxes, out.counts, self.roi_size, upsample)
                    done[lane] = torch.cuda.Event()
                    done[lane].record(side)
            for s_ in self._side:
                cur.wait_stream(s_)
        return bld.finish()

    def run(self, provider, ordered: bool = True, proto_provider=None,
            mask_words_per_row: float = 56.0) -> Dict[str, torch.Tensor]:
        """detect -> merge -> masks.  (Ordering the survivors on a side stream beside the mask pass was measured: the
        sort fills the SMs whenever it runs, 72.6 -> 73.3 ms per slide; the phases stay in sequence.)"""
        self.detect(provider, keep_batches=proto_provider is not None)
        res = self.merge(ordered)
        if proto_provider is not None:
            res['masks'] = self.masks(proto_provider, res['state'], words_per_row=mask_words_per_row)
        return res
