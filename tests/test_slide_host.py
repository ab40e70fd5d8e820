"""CPU tests of the host-side slide logic: the tile scanner against the reference golden, tile cores."""
import torch

from conftest import load_golden
from hd_yolo_b200 import slide as hs
from oracle import port


def test_sliding_window_scanner_vs_reference_golden():
    g = load_golden("tile_merge")
    H, W = g["image_size"].tolist()
    assert torch.equal(hs.sliding_window_scanner((H, W), tuple(g["roi_size"].tolist()), int(g["overlap"])),
                       torch.from_numpy(g["rois"]))
    big = hs.sliding_window_scanner((100000, 100000), (1024, 1024), 64)
    assert len(big) == int(g["scan_slide_n"]) == 11025
    assert torch.equal(big[:3], torch.from_numpy(g["scan_slide_head"]))
    assert torch.equal(big[-3:], torch.from_numpy(g["scan_slide_tail"]))
    assert torch.equal(hs.sliding_window_scanner((300, 500), (128, 200), 0), torch.from_numpy(g["scan_small"]))
    assert torch.equal(hs.sliding_window_scanner((100, 100), (128, 128), 16), torch.from_numpy(g["scan_fit"]))
    assert hs.sliding_window_scanner((30, 50)).tolist() == [[0., 0., 30., 50.]]
    for args in [((700, 900), 256, 64), ((512, 512), (128, 256), 0), ((1000, 300), (256, 512), 100)]:
        assert torch.equal(hs.sliding_window_scanner(*args), port.sliding_window_scanner(*args))


def test_tile_cores_are_untouched_by_other_tiles():
    rois = hs.sliding_window_scanner((700, 900), (256, 256), 64)
    cores = hs.tile_cores(rois)
    for i, c in enumerate(cores):
        x0, y0, x1, y1 = [max(min(v, 1e6), -1e6) for v in c.tolist()]
        if x1 <= x0 or y1 <= y0:
            continue
        for j, r in enumerate(rois):
            if i == j:
                continue
            # open-interval intersection of the core with any other tile is empty
            ix = min(x1, float(r[2])) - max(x0, float(r[0]))
            iy = min(y1, float(r[3])) - max(y0, float(r[1]))
            assert ix <= 0 or iy <= 0, (i, j)
    # interior tiles of a 256/64 tiling keep a 128-px core
    inner = cores[6]                                   # row 1, column 1 of the 5 x 4 grid
    assert (inner[2] - inner[0]).item() == 128.0 and (inner[3] - inner[1]).item() == 128.0
