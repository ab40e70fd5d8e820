"""Host-side mirror of the reference's mask post-processing over the C-ABI kernels (csrc/mask.cu).

Same names, argument meaning and return structures as the functions they replace:

* ``mask_select``            -- mask tail of Detect.compute_outputs (metayolo/models/yolo_head.py:332, 346-353)
* ``paste_masks_in_image``   -- torchvision.models.detection.roi_heads.paste_masks_in_image as the reference calls
                                it (metayolo/val_nuclei.py:169-176, metayolo/models/evaluation.py:122-123)
* ``process_mask``           -- ultralytics/yolov5 v7 utils/segment/general.py::process_mask (north-star extension;
                                the reference does not vendor it)

plus the bit-packed variants the throughput numbers use (``paste_masks_packed``, ``process_mask_packed``), which
fuse the ``> 0.5`` threshold (M3) and write box-cropped bit planes instead of [k,H,W] fp32 canvases.
No CPU fallback: CPU tensors raise HdyError.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import HdyError, ptr
from .ops import _Scratch, _aligned16, _call, _need_cuda, _need_head_tensor, _stream

_pm_scratch = _Scratch()

__all__ = ["PackedMasks", "mask_select", "paste_masks_in_image", "paste_masks_packed", "process_mask",
           "process_mask_batch", "process_mask_packed", "SlideMaskBuilder"]


@dataclass
class PackedMasks:
    """Box-cropped bit planes (include/hd_yolo_b200.h, "Cropped bit-packed layout").

    geom[i] = (x0, y0, w, h): the window of mask i inside the (H, W) canvas; offsets[i] = first 32-bit word of
    mask i inside ``bits``; row r of mask i is ceil(w/32) words, pixel (x, y) is bit (x-x0)&31 of word
    offsets[i] + (y-y0)*ceil(w/32) + ((x-x0)>>5)."""

    geom: torch.Tensor      # [K, 4] int32
    offsets: torch.Tensor   # [K+1] int64
    bits: torch.Tensor      # [words] int32 (bit pattern of uint32)
    H: int
    W: int
    status: torch.Tensor    # [1] int32 device status word (HDY_STATUS_OVERFLOW)

    def __len__(self):
        return self.geom.shape[0]

    def check(self) -> None:
        if int(self.status.item()) & _lib.HDY_STATUS_OVERFLOW:
            raise HdyError("bit-plane capacity overflow: pass a larger capacity_words")

    def to_dense(self) -> torch.Tensor:
        """[K, H, W] uint8 canvas of 0/1 (what ``paste_masks_in_image(...) > 0.5`` holds)."""
        K = len(self)
        out = torch.empty((K, self.H, self.W), dtype=torch.uint8, device=self.geom.device)
        if K:
            _call("hdy_unpack_masks", ptr(self.geom), ptr(self.offsets), ptr(self.bits), K, self.H, self.W, ptr(out),
                  _stream(), launches=2)
        return out

    def nbytes(self) -> int:
        return self.bits.numel() * 4 + self.geom.numel() * 4 + self.offsets.numel() * 8


def _hw(shape) -> Tuple[int, int]:
    if isinstance(shape, int):
        return int(shape), int(shape)
    return int(shape[0]), int(shape[1])


def _boxes4(boxes: torch.Tensor, k: int, name="boxes") -> torch.Tensor:
    _need_cuda(boxes, name)
    if boxes.dim() != 2 or boxes.shape[1] < 4 or boxes.shape[0] != k:
        raise HdyError(f"{name} must be [{k}, 4], got {tuple(boxes.shape)}")
    b = boxes[:, :4]
    return _aligned16(b.contiguous())


# ------------------------------------------------------------------------------------------------ M1
def mask_select(mask_logits: torch.Tensor, labels: torch.Tensor, mask_indices: torch.Tensor) -> torch.Tensor:
    """yolo_head.py:332, 346-353 for the K detections of one batch: sigmoid, pick the channel
    ``mask_indices[labels.clamp(min=0)]`` of every detection, zero masks whose channel index is < 0.
    mask_logits [K, C, M, M] fp32, labels [K] int64, mask_indices [1+nc] int64 -> [K, 1, M, M]."""
    _need_cuda(mask_logits, "mask_logits")
    if mask_logits.dim() != 4 or mask_logits.shape[2] != mask_logits.shape[3]:
        raise HdyError(f"mask_logits must be [K, C, M, M], got {tuple(mask_logits.shape)}")
    K, Cn, M, _ = mask_logits.shape
    dev = mask_logits.device
    labels = labels.to(dev, torch.int64).contiguous()
    mask_indices = mask_indices.to(dev, torch.int64).contiguous()
    if labels.shape != (K,):
        raise HdyError("labels must be [K]")
    out = torch.empty((K, 1, M, M), dtype=torch.float32, device=dev)
    if K:
        _call("hdy_mask_select", ptr(mask_logits.contiguous()), ptr(labels), ptr(mask_indices), K, Cn, M, ptr(out),
              _stream())
    return out


# ------------------------------------------------------------------------------------------------ M2
def _paste_src(masks: torch.Tensor):
    _need_cuda(masks, "masks")
    if masks.dim() == 3:
        masks = masks[:, None]
    if masks.dim() != 4 or masks.shape[2] != masks.shape[3]:
        raise HdyError(f"masks must be [k, C, M, M], got {tuple(masks.shape)}")
    return masks.contiguous()


def paste_masks_in_image(masks: torch.Tensor, boxes: torch.Tensor, img_shape, padding: int = 1) -> torch.Tensor:
    """torchvision paste_masks_in_image(masks [k,1,M,M], boxes [k,4], (H,W), padding=1) -> [k,1,H,W] fp32
    (val_nuclei.py:169-176, evaluation.py:122-123): zero-pad the mask, grow the box by (M+2p)/M about its
    centre, truncate to integers, bilinear-resize to the box, paste into a zero canvas."""
    masks = _paste_src(masks)
    k, Cn, M, _ = masks.shape
    if Cn != 1:
        raise HdyError("paste_masks_in_image takes [k,1,M,M] masks (use paste_masks_packed(channel=...) to fuse M1)")
    H, W = _hw(img_shape)
    out = torch.empty((k, 1, H, W), dtype=torch.float32, device=masks.device)
    if k:
        b = _boxes4(boxes, k)
        _call("hdy_paste_masks", ptr(masks), None, ptr(b), k, 1, M, int(padding), 0, H, W, ptr(out), _stream(),
              launches=2)
    return out


def paste_masks_packed(masks: torch.Tensor, boxes: torch.Tensor, img_shape, padding: int = 1,
                       channel: Optional[torch.Tensor] = None, apply_sigmoid: bool = False,
                       capacity_words: Optional[int] = None) -> PackedMasks:
    """M1 + M2 + M3 fused: ``paste_masks_in_image(...) > 0.5`` as box-cropped bit planes.

    masks [k, C, M, M]: probabilities, or raw logits with apply_sigmoid=True; channel [k] int32 selects the
    channel per detection (None: 0; < 0: empty mask) -- i.e. ``mask_indices[labels.clamp(min=0)]``.
    capacity_words=None sizes ``bits`` exactly (one 8-byte device->host read); pass an upper bound to stay
    asynchronous (overflow is then reported by PackedMasks.check())."""
    masks = _paste_src(masks)
    k, Cn, M, _ = masks.shape
    H, W = _hw(img_shape)
    dev = masks.device
    geom = torch.empty((k, 4), dtype=torch.int32, device=dev)
    offsets = torch.empty((k + 1,), dtype=torch.int64, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    b = _boxes4(boxes, k) if k else None
    if channel is not None:
        channel = channel.to(dev, torch.int32).contiguous()
    _call("hdy_paste_geometry", ptr(b), ptr(channel), k, M, int(padding), H, W, ptr(geom), ptr(offsets), _stream(),
          launches=4)
    words = int(offsets[k].item()) if capacity_words is None else int(capacity_words)
    bits = torch.empty((max(words, 1),), dtype=torch.int32, device=dev)
    if k:
        _call("hdy_paste_masks_packed", ptr(masks), ptr(channel), ptr(b), ptr(offsets), k, Cn, M, int(padding),
              int(bool(apply_sigmoid)), H, W, ptr(bits), words, ptr(status), _stream())
    return PackedMasks(geom, offsets, bits[:words], H, W, status)


# ---------------------------------------------------------------------------------- process_mask (B)
def _pm_args(protos, coef, boxes, counts):
    _need_head_tensor(protos, "protos")
    _need_cuda(coef, "masks_in")
    _need_cuda(boxes, "bboxes")
    if protos.dim() != 4 or coef.dim() != 3 or boxes.dim() != 3:
        raise HdyError("expected protos [bs,nm,mh,mw], coef [bs,max_det,nm], boxes [bs,max_det,4]")
    bs, nm, mh, mw = protos.shape
    if coef.shape[0] != bs or coef.shape[2] != nm or boxes.shape[:2] != coef.shape[:2] or boxes.shape[2] != 4:
        raise HdyError("protos / coef / boxes shapes disagree")
    if counts.shape != (bs,) or counts.dtype != torch.int32 or not counts.is_cuda:
        raise HdyError("counts must be a CUDA int32 tensor [bs]")
    return bs, nm, mh, mw, coef.shape[1]


def _pm_workspace(dev, bs: int, md: int):
    """Grow-only scratch for the two-phase path (sigmoid patches + the list of over-sized detections)."""
    nbytes = _lib.load().hdy_process_mask_workspace_bytes(bs, md)
    return _pm_scratch.get(dev, "process_mask", nbytes), nbytes


def process_mask_batch(protos: torch.Tensor, coef: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, shape,
                       upsample: bool = False) -> torch.Tensor:
    """process_mask for a batch of tiles (struct-of-arrays, as DetectBatch holds them): protos [bs,nm,mh,mw],
    coef [bs,max_det,nm], boxes [bs,max_det,4] in image pixels, counts [bs] int32.
    Returns [bs, max_det, oh, ow] fp32 in {0,1}; slots >= counts[i] are zero."""
    bs, nm, mh, mw, md = _pm_args(protos, coef, boxes, counts)
    ih, iw = _hw(shape)
    oh, ow = (ih, iw) if upsample else (mh, mw)
    out = torch.empty((bs, md, oh, ow), dtype=torch.float32, device=protos.device)
    coef = coef.float()
    if bs and md:
        ws, wbytes = _pm_workspace(protos.device, bs, md)
        _call("hdy_process_mask", ptr(_aligned16(protos.contiguous())), _need_head_tensor(protos, "protos"),
              ptr(coef.contiguous()),
              ptr(_aligned16(boxes.contiguous())), ptr(counts), bs, md, nm, mh, mw, ih, iw, int(bool(upsample)),
              ptr(out), ptr(ws), wbytes, _stream(), launches=4)
    return out


def process_mask(protos: torch.Tensor, masks_in: torch.Tensor, bboxes: torch.Tensor, shape,
                 upsample: bool = False) -> torch.Tensor:
    """ultralytics/yolov5 v7 ``process_mask(protos [c,mh,mw], masks_in [n,c], bboxes [n,4], shape (ih,iw),
    upsample=False)`` -> [n, mh, mw] (or [n, ih, iw] with upsample) of 0/1 floats:
    sigmoid(masks_in @ protos), crop to the box scaled into proto space, optional bilinear upsample, > 0.5."""
    _need_head_tensor(protos, "protos")
    if protos.dim() != 3:
        raise HdyError("protos must be [c, mh, mw]")
    n = masks_in.shape[0]
    ih, iw = _hw(shape)
    if n == 0:
        oh, ow = (ih, iw) if upsample else tuple(protos.shape[1:])
        return protos.new_zeros((0, oh, ow), dtype=torch.float32)
    counts = torch.full((1,), n, dtype=torch.int32, device=protos.device)
    return process_mask_batch(protos[None], masks_in.float()[None], bboxes.float()[None, :, :4], counts, shape,
                              upsample)[0]


def process_mask_packed(protos: torch.Tensor, coef: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, shape,
                        upsample: bool = False, capacity_words: Optional[int] = None,
                        row_state: Optional[torch.Tensor] = None,
                        tile_offsets: Optional[torch.Tensor] = None) -> PackedMasks:
    """Bit-packed process_mask over a batch of tiles; mask index = tile * max_det + slot (empty for slots
    >= counts[tile]).  Window coordinates are in output pixels ((ih,iw) with upsample, else (mh,mw)).
    row_state [n] uint8 + tile_offsets [bs+1] int64 (slide form): only slots whose slide row was KEPT get a mask."""
    bs, nm, mh, mw, md = _pm_args(protos, coef, boxes, counts)
    ih, iw = _hw(shape)
    dev = protos.device
    K = bs * md
    boxes = _aligned16(boxes.contiguous())
    geom = torch.empty((K, 4), dtype=torch.int32, device=dev)
    offsets = torch.empty((K + 1,), dtype=torch.int64, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    _call("hdy_process_mask_geometry", ptr(boxes), ptr(counts), bs, md, mh, mw, ih, iw, int(bool(upsample)),
          ptr(row_state), ptr(tile_offsets), ptr(geom), ptr(offsets), _stream(), launches=2)
    words = int(offsets[K].item()) if capacity_words is None else int(capacity_words)
    bits = torch.empty((max(words, 1),), dtype=torch.int32, device=dev)
    if K:
        ws, wbytes = _pm_workspace(dev, bs, md)
        _call("hdy_process_mask_packed", ptr(_aligned16(protos.contiguous())), _need_head_tensor(protos, "protos"),
              ptr(coef.contiguous()), ptr(boxes),
              ptr(counts), ptr(geom), ptr(offsets), bs, md, nm, mh, mw, ih, iw, int(bool(upsample)), ptr(bits), words,
              ptr(status),
              ptr(ws), wbytes, _stream(), launches=3)
    oh, ow = (ih, iw) if upsample else (mh, mw)
    return PackedMasks(geom, offsets, bits[:words], oh, ow, status)


class SlideMaskBuilder:
    """Masks of a whole slide, computed AFTER the slide-level verdicts and only for the rows Ensemble.merge keeps
    (yolo.py:197-202 gathers masks[keep]; the masks of suppressed duplicates are never materialised).  One
    PackedMasks over the slide's rows: geom[r] is the window of row r in SLIDE pixels, offsets[r] its first word in the
    slide-wide `bits`; rows that were not kept have empty windows.  Everything stays on the device; `check()` reads
    the overflow flag."""

    def __init__(self, n_rows: int, capacity_words: int, image_size, device):
        d = torch.device(device)
        self.n = int(n_rows)
        self.capacity_words = int(capacity_words)
        self.geom = torch.zeros((max(self.n, 1), 4), dtype=torch.int32, device=d)
        self.offsets = torch.zeros((self.n + 1,), dtype=torch.int64, device=d)
        self.bits = torch.empty((max(self.capacity_words, 1),), dtype=torch.int32, device=d)
        self.cursor2 = torch.zeros((2,), dtype=torch.int64, device=d)
        self.status = torch.zeros((1,), dtype=torch.int32, device=d)
        self.batches = 0
        self.H, self.W = _hw(image_size)

    def prepare(self, boxes: torch.Tensor, counts: torch.Tensor, tile_offsets: torch.Tensor, rois: torch.Tensor,
                row_state: torch.Tensor, proto_hw, shape, upsample: bool = True, lane: int = 0):
        """Windows and word offsets of the next batch (in batch order: the slide-wide word cursor advances), scattered
        to the slide rows.  Returns (geom, offsets) for ``run``; `lane` picks the scratch pair, so that the kernels of
        batch i can still read theirs while batch i + 1 is prepared (mask kernels on alternating streams)."""
        bs, md = int(boxes.shape[0]), int(boxes.shape[1])
        mh, mw = proto_hw
        ih, iw = _hw(shape)
        dev = boxes.device
        K = bs * md
        if K == 0:
            return None
        boxes = _aligned16(boxes.contiguous())
        geom = _pm_scratch.get(dev, f"slide_geom{lane}", K * 16).view(torch.int32)[:K * 4]
        offsets = _pm_scratch.get(dev, f"slide_offsets{lane}", (K + 1) * 8).view(torch.int64)[:K + 1]
        _call("hdy_process_mask_geometry", ptr(boxes), ptr(counts), bs, md, mh, mw, ih, iw, int(bool(upsample)),
              ptr(row_state), ptr(tile_offsets), ptr(geom), ptr(offsets), _stream(), launches=2)
        _call("hdy_process_mask_rows", ptr(geom), ptr(offsets), ptr(counts), ptr(tile_offsets),
              ptr(_aligned16(rois.contiguous())), bs, md, ptr(self.cursor2), self.batches & 1, ptr(self.geom),
              ptr(self.offsets), _stream())
        self.batches += 1
        return geom, offsets

    def run(self, prepared, protos: torch.Tensor, coef: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor,
            shape, upsample: bool = True) -> None:
        """The mask kernels of a prepared batch (any stream ordered after its ``prepare``)."""
        if prepared is None:
            return
        geom, offsets = prepared
        bs, nm, mh, mw, md = _pm_args(protos, coef, boxes, counts)
        ih, iw = _hw(shape)
        dev = protos.device
        boxes = _aligned16(boxes.contiguous())
        ws, wbytes = _pm_workspace(dev, bs, md)
        _call("hdy_process_mask_packed", ptr(_aligned16(protos.contiguous())), _need_head_tensor(protos, "protos"),
              ptr(coef.contiguous()), ptr(boxes),
              ptr(counts), ptr(geom), ptr(offsets), bs, md, nm, mh, mw, ih, iw, int(bool(upsample)), ptr(self.bits),
              self.capacity_words, ptr(self.status), ptr(ws), wbytes, _stream(), launches=3)

    def add_batch(self, protos: torch.Tensor, coef: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor,
                  tile_offsets: torch.Tensor, rois: torch.Tensor, row_state: torch.Tensor, shape,
                  upsample: bool = True) -> None:
        """protos [bs,nm,mh,mw] of the batch's tiles; coef / boxes / counts as DetectBatch holds them (tile
        coordinates); tile_offsets [bs+1] from the append; rois [bs,4]; row_state: verdict per slide row."""
        prep = self.prepare(boxes, counts, tile_offsets, rois, row_state, tuple(protos.shape[2:]), shape, upsample)
        self.run(prep, protos, coef, boxes, counts, shape, upsample)

    def finish(self) -> PackedMasks:
        """offsets[n] = total words.  Rows that were never live keep offset 0 / an empty window: consumers address a
        mask through geom[r] and offsets[r] only."""
        self.offsets[self.n:self.n + 1].copy_(self.cursor2[self.batches & 1:(self.batches & 1) + 1])
        return PackedMasks(self.geom[:self.n], self.offsets, self.bits, self.H, self.W, self.status)
