"""Compact on-disk format for post-processing results (the reference pickles Python dicts of CPU
tensors with ``torch.save``, evaluation.py:225 -- 4 MB of dense fp32 mask per nucleus at 1024 px if masks are kept).

One little-endian file: a JSON header line (magic, version, array table with dtype / shape / byte offset, free-form
``meta``) padded to 64 bytes, then the raw arrays, each 64-byte aligned.  Detections are the struct-of-arrays the
package works in (``boxes`` fp32 [k,4], ``scores`` fp32 [k], ``labels`` int64 stored as int16, optional ``index``);
masks are the cropped bit planes of ``PackedMasks`` (geom / offsets / bits), ~100-300 bytes per nucleus.  Host-side
only (CUDA tensors are read back once); loading memory-maps the file and returns CPU tensors.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from ._lib import HdyError
from .masks import PackedMasks

__all__ = ["save_detections", "load_detections"]

MAGIC = "hd_yolo_b200.detections"
VERSION = 1
_ALIGN = 64
# label -100 ("unclassified", yolo_head.py:345) and class ids up to 32767 fit int16
_STORE_AS = {"labels": np.int16}


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().numpy()


def save_detections(path: str, result: Dict[str, torch.Tensor], masks: Optional[PackedMasks] = None,
                    meta: Optional[dict] = None) -> int:
    """Writes `result` ('boxes' [k,4], 'scores' [k], optional 'labels' [k], 'index' [k], any other tensor entry) and
    optionally the bit-packed masks of the same k detections.  Returns the file size in bytes."""
    arrays = {}
    for name, t in result.items():
        if not isinstance(t, torch.Tensor):
            continue
        a = _np(t)
        if name in _STORE_AS:
            if a.size and (a.min() < np.iinfo(_STORE_AS[name]).min or a.max() > np.iinfo(_STORE_AS[name]).max):
                raise HdyError(f"'{name}' does not fit {_STORE_AS[name].__name__}")
            a = a.astype(_STORE_AS[name])
        arrays[name] = (a, str(t.dtype).replace("torch.", ""))
    k = int(result["boxes"].shape[0])
    if masks is not None:
        if len(masks) != k:
            raise HdyError(f"{len(masks)} masks for {k} detections")
        masks.check()
        words = int(masks.offsets[-1].item())
        arrays["mask_geom"] = (_np(masks.geom), "int32")
        arrays["mask_offsets"] = (_np(masks.offsets), "int64")
        arrays["mask_bits"] = (_np(masks.bits[:words]), "int32")
    table, off = [], 0
    for name, (a, orig) in arrays.items():
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        table.append({"name": name, "dtype": a.dtype.str, "shape": list(a.shape), "offset": off, "restore": orig})
        off += a.nbytes
    header = {"magic": MAGIC, "version": VERSION, "count": k, "arrays": table, "meta": meta or {},
              "mask_canvas": [masks.H, masks.W] if masks is not None else None}
    head = (json.dumps(header) + "\n").encode()
    head += b" " * ((-len(head)) % _ALIGN)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(head)
        for entry, (a, _) in zip(table, arrays.values()):
            f.seek(len(head) + entry["offset"])
            f.write(a.tobytes())
    os.replace(tmp, path)          # a reader never sees a half-written file
    return os.path.getsize(path)


def load_detections(path: str, mmap: bool = True) -> Tuple[Dict[str, torch.Tensor], Optional[PackedMasks], dict]:
    """-> (result dict of CPU tensors, PackedMasks on the CPU or None, meta).  `PackedMasks.to_dense()` needs the
    tensors on a CUDA device (`.geom.cuda()` ...): the file itself never holds dense masks."""
    with open(path, "rb") as f:
        line = f.readline()
    try:
        header = json.loads(line)
    except Exception as e:
        raise HdyError(f"{path}: not a detections file") from e
    if header.get("magic") != MAGIC or header.get("version") != VERSION:
        raise HdyError(f"{path}: magic/version mismatch ({header.get('magic')}, {header.get('version')})")
    base = (len(line) + _ALIGN - 1) // _ALIGN * _ALIGN
    out: Dict[str, torch.Tensor] = {}
    for e in header["arrays"]:
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if mmap and n:
            a = np.memmap(path, dtype=np.dtype(e["dtype"]), mode="r", offset=base + e["offset"], shape=tuple(e["shape"]))
            a = np.array(a)        # one sequential read; the tensors own their memory
        else:
            a = np.fromfile(path, dtype=np.dtype(e["dtype"]), count=n, offset=base + e["offset"]).reshape(e["shape"])
        out[e["name"]] = torch.from_numpy(np.ascontiguousarray(a)).to(getattr(torch, e["restore"]))
    masks = None
    if header.get("mask_canvas"):
        H, W = header["mask_canvas"]
        masks = PackedMasks(out.pop("mask_geom"), out.pop("mask_offsets"), out.pop("mask_bits"), int(H), int(W),
                            torch.zeros((1,), dtype=torch.int32))
    return out, masks, header.get("meta", {})
