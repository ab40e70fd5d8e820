"""CPU tests of the compact detections file (hd_yolo_b200/io.py): round trip of detections and bit-packed masks,
alignment of the arrays, atomic replace, rejection of foreign files."""
import json
import os

import pytest
import torch

import hd_yolo_b200 as hdy
from hd_yolo_b200.masks import PackedMasks


def _result(k, seed=0):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand((k, 2), generator=g) * 1e5
    s = torch.rand((k, 2), generator=g) * 30 + 8
    labels = torch.randint(1, 5, (k,), generator=g)
    labels[::7] = -100                                    # "unclassified" (yolo_head.py:345)
    return {'boxes': torch.cat([c - s / 2, c + s / 2], 1), 'scores': torch.rand(k, generator=g), 'labels': labels,
            'index': torch.randperm(k, generator=g)}


def _masks(k, seed=1):
    g = torch.Generator().manual_seed(seed)
    geom = torch.stack([torch.randint(0, 900, (k,), generator=g), torch.randint(0, 900, (k,), generator=g),
                        torch.randint(1, 40, (k,), generator=g), torch.randint(1, 40, (k,), generator=g)], 1).int()
    words = ((geom[:, 2] + 31) // 32).long() * geom[:, 3].long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(words, 0)])
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (int(offsets[-1]) + 17,), generator=g, dtype=torch.int64).int()
    return PackedMasks(geom, offsets, bits, 1024, 1024, torch.zeros(1, dtype=torch.int32))


@pytest.mark.parametrize("k", [0, 1, 1000])
def test_round_trip(tmp_path, k):
    res, pm = _result(k), _masks(k)
    path = str(tmp_path / "slide.det")
    size = hdy.save_detections(path, res, pm, meta={"slide": "S1", "conf": 0.25})
    assert size == os.path.getsize(path) and not os.path.exists(path + ".tmp")
    out, masks, meta = hdy.load_detections(path)
    assert meta == {"slide": "S1", "conf": 0.25}
    for name in res:
        assert out[name].dtype == res[name].dtype and torch.equal(out[name], res[name]), name
    words = int(pm.offsets[-1])
    assert masks.H == 1024 and torch.equal(masks.geom, pm.geom) and torch.equal(masks.offsets, pm.offsets)
    assert torch.equal(masks.bits, pm.bits[:words])       # capacity slack is not written
    if k == 1000:                                         # 16 + 4 + 2 + 8 bytes per detection + ~150 per mask
        assert size < k * (30 + 16 + 8) + 4 * words + 4096
    out2, _, _ = hdy.load_detections(path, mmap=False)
    assert all(torch.equal(out2[n], res[n]) for n in res)


def test_arrays_are_aligned_and_described(tmp_path):
    path = str(tmp_path / "a.det")
    hdy.save_detections(path, _result(10))
    with open(path, "rb") as f:
        line = f.readline()
    header = json.loads(line)
    assert header["magic"] == "hd_yolo_b200.detections" and header["count"] == 10 and header["mask_canvas"] is None
    assert all(e["offset"] % 64 == 0 for e in header["arrays"])
    assert {e["name"]: e["dtype"] for e in header["arrays"]}["labels"] == "<i2"
    out, masks, _ = hdy.load_detections(path)
    assert masks is None and out["labels"].dtype == torch.int64


def test_rejects_foreign_files_and_bad_labels(tmp_path):
    path = str(tmp_path / "x.det")
    with open(path, "w") as f:
        f.write("not a detections file\n")
    with pytest.raises(hdy.HdyError):
        hdy.load_detections(path)
    res = _result(4)
    res['labels'][0] = 70000
    with pytest.raises(hdy.HdyError):
        hdy.save_detections(path, res)
    with pytest.raises(hdy.HdyError):
        hdy.save_detections(path, _result(4), _masks(5))
