"""GPU parity tests of the tile merge (T2/T3), coordinates (C1) and the radix sort (pytest -m gpu), through the C-ABI,
against the reference golden (tests/golden/tile_merge.npz) and the oracle port.  Kept sets, order and labels are
bit-exact; merged boxes are bit-exact (one fp32 addition per coordinate)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, unpack_list
import hd_yolo_b200 as hdy
from hd_yolo_b200 import slide as hs
from hd_yolo_b200.ops import DetectBatch
from oracle import nms_c, port

pytestmark = pytest.mark.gpu


def _to(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) and k != 'roi' else v) for k, v in d.items()}


def test_merge_outputs_and_ensemble_vs_reference_golden(cuda_device):
    g = load_golden("tile_merge")
    rois = torch.from_numpy(g["rois"])
    tiles = unpack_list(g, "tile", ["boxes", "scores", "labels"])
    for t, r in zip(tiles, rois):
        t["roi"] = r
    merged = hs.merge_outputs([_to(t, cuda_device) for t in tiles])
    for k in ("boxes", "scores", "labels"):
        assert torch.equal(merged[k].cpu(), torch.from_numpy(g["merged_" + k])), k
    conf, iou, md = g["params"].tolist()
    ens = hs.Ensemble([], nms_params={'conf_thres': conf, 'iou_thres': iou, 'max_det': md})
    final = ens.merge([{'det': merged}])['det']
    for k in ("boxes", "scores", "labels"):
        assert torch.equal(final[k].cpu(), torch.from_numpy(g["final_" + k])), k
    assert final['labels'].dtype == torch.int64
    # max_det cap keeps the best-scored prefix
    ens2 = hs.Ensemble([], nms_params={'conf_thres': conf, 'iou_thres': iou, 'max_det': 37})
    f2 = ens2.merge([{'det': merged}])['det']
    assert torch.equal(f2['boxes'].cpu(), torch.from_numpy(g["final_boxes"])[:37])


def _synthetic_slide(seed, H, W, roi, overlap, n_nuclei, jitter=1.0, coord_offset=0.0):
    """Per-tile detection lists from a global nuclei field: neighbouring tiles re-detect the nuclei in their overlap
    band with <= jitter px box noise and their own scores (SURVEY 8d cfg 4)."""
    g = torch.Generator().manual_seed(seed)
    rois = hs.sliding_window_scanner((H, W), (roi, roi), overlap)
    nuc = torch.rand((n_nuclei, 2), generator=g) * torch.tensor([W, H], dtype=torch.float32)
    size = torch.rand((n_nuclei, 2), generator=g) * 24 + 12
    tiles = []
    for r in rois:
        x0, y0, x1, y1 = r.tolist()
        inside = (nuc[:, 0] > x0 + 2) & (nuc[:, 0] < x1 - 2) & (nuc[:, 1] > y0 + 2) & (nuc[:, 1] < y1 - 2)
        k = int(inside.sum())
        c = nuc[inside] - torch.tensor([x0, y0]) + (torch.rand((k, 2), generator=g) - 0.5) * jitter
        s = size[inside] + (torch.rand((k, 2), generator=g) - 0.5) * jitter
        boxes = torch.cat([c - s / 2, c + s / 2], 1)
        scores = ((torch.rand(k, generator=g) * 0.85 + 0.1) * 256).round() / 256     # ties across tiles
        labels = torch.randint(1, 5, (k,), generator=g)
        roi_shift = r + coord_offset
        tiles.append({'boxes': boxes, 'scores': scores, 'labels': labels, 'roi': roi_shift})
    return rois, tiles


def _as_batch(tiles, dev):
    bs, md = len(tiles), max(max(len(t['boxes']) for t in tiles), 1)
    boxes = torch.zeros((bs, md, 4))
    scores = torch.zeros((bs, md))
    labels = torch.zeros((bs, md), dtype=torch.int64)
    counts = torch.zeros((bs,), dtype=torch.int32)
    for i, t in enumerate(tiles):
        k = len(t['boxes'])
        boxes[i, :k], scores[i, :k], labels[i, :k], counts[i] = t['boxes'], t['scores'], t['labels'], k
    z = torch.zeros((bs, md), device=dev)
    return DetectBatch(boxes.to(dev), None, scores.to(dev), labels.to(dev), z, None, z.int(), counts.to(dev),
                       torch.zeros(bs + 1, dtype=torch.int32, device=dev), md)


@pytest.mark.parametrize("seed,H,W,roi,overlap,n,off", [(1, 900, 1300, 256, 64, 2500, 0.0), (2, 2000, 2000, 512, 64, 9000, 0.0),
                                                        (3, 1500, 1100, 256, 32, 4000, 98304.0), (4, 600, 600, 1024, 64, 500, 0.0)])
def test_slide_accumulator_vs_oracle(cuda_device, seed, H, W, roi, overlap, n, off):
    rois, tiles = _synthetic_slide(seed, H, W, roi, overlap, n, coord_offset=off)
    params = {'conf_thres': 0.2, 'iou_thres': 0.45, 'max_det': 10 ** 9}
    ref_m = port.merge_outputs(tiles)
    ref = port.ensemble_merge([{'det': ref_m}], params)['det']
    acc = hs.SlideAccumulator(capacity=sum(len(t['boxes']) for t in tiles) + 100, device=cuda_device)
    half = max(1, len(tiles) // 2)
    for chunk in (tiles[:half], tiles[half:]):           # two batches: offsets carry over on the device
        if chunk:
            acc.append(_as_batch(chunk, cuda_device), torch.stack([t['roi'] for t in chunk]).to(cuda_device))
    assert acc.count() == len(ref_m['boxes'])
    n_rows = acc.count()
    assert torch.equal(acc.boxes[:n_rows].cpu(), ref_m['boxes'])
    out = acc.merge(0.2, 0.45)
    assert len(ref['boxes']) < n_rows
    assert torch.equal(out['boxes'].cpu(), ref['boxes'])
    assert torch.equal(out['scores'].cpu(), ref['scores'])
    assert torch.equal(out['labels'].cpu(), ref['labels'])
    # 'index' addresses the concatenated order
    assert torch.equal(ref_m['boxes'][out['index'].cpu()], ref['boxes'])
    # the interior shortcut is refused without the gray-zone flags of the per-tile NMS (these batches have none);
    # its exactness with flags is covered by test_interior_shortcut_is_exact_far_from_the_origin
    with pytest.raises(hdy.HdyError, match="gray-zone"):
        acc.verdicts(0.2, 0.45, interior_shortcut=True)


def test_merge_nms_random_dense_vs_oracle(cuda_device):
    g = torch.Generator().manual_seed(5)
    for n, span, lo, hi in [(1, 10, 4, 8), (5000, 700, 8, 40), (40000, 3000, 10, 36), (3000, 200, 5, 300)]:
        c = torch.rand((n, 2), generator=g) * span
        wh = torch.rand((n, 2), generator=g) * (hi - lo) + lo
        b = torch.cat([c - wh / 2, c + wh / 2], 1)
        s = (torch.rand(n, generator=g) * 500).round() / 500
        keep_ref = nms_c.nms(b[s > 0.1].numpy(), s[s > 0.1].numpy(), 0.45)
        idx_ref = torch.nonzero(s > 0.1)[:, 0][torch.from_numpy(keep_ref)]
        st = hs.merge_nms(b.to(cuda_device), s.to(cuda_device), 0.1, 0.45).cpu()
        assert set(torch.nonzero(st == hs.STATE_KEPT)[:, 0].tolist()) == set(idx_ref.tolist())
        assert bool(((st == hs.STATE_DROPPED) == (s <= np.float32(0.1))).all())
        assert not bool((st == 0).any())


def test_merge_empty_and_all_dropped(cuda_device):
    assert hs.merge_nms(torch.zeros((0, 4), device=cuda_device), torch.zeros(0, device=cuda_device), 0.2, 0.45).numel() == 0
    b = torch.tensor([[0., 0., 10., 10.], [1., 1., 11., 11.]], device=cuda_device)
    s = torch.tensor([0.1, 0.15], device=cuda_device)
    assert hs.merge_nms(b, s, 0.2, 0.45).tolist() == [3, 3]
    out = hs.ensemble_merge([{'t': {'boxes': b, 'scores': s, 'labels': torch.tensor([1, 2], device=cuda_device)}}],
                            {'conf_thres': 0.2, 'iou_thres': 0.45, 'max_det': 10})['t']
    assert out['boxes'].shape == (0, 4) and out['labels'].shape == (0,)


@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 50000, 1 << 20])
def test_radix_sort(cuda_device, n):
    g = torch.Generator().manual_seed(n)
    k = torch.randint(-2 ** 62, 2 ** 62, (n,), generator=g, dtype=torch.int64)
    k[::7] = k[0]                                           # duplicates
    out = hs.sort_keys(k.to(cuda_device).clone()).cpu()
    as_u = lambda t: t.numpy().view(np.uint64)
    assert np.array_equal(as_u(out), np.sort(as_u(k)))


def test_scale_clip_rescale_vs_oracle(cuda_device):
    g = torch.Generator().manual_seed(8)
    coords = torch.rand((500, 6), generator=g) * 700 - 30
    for img1, img0, rp in [((640, 640), (480, 720), None), (640, (1000, 1000), None), ((384, 640), (1080, 1920), None),
                           ((640, 640), (480, 720), ((0.8, 0.8), (13.0, 40.0)))]:
        ref = port.scale_coords(img1, coords.clone(), img0, rp)
        out = coords.clone().to(cuda_device)
        ret = hs.scale_coords(img1, out, img0, rp)
        assert ret is out and torch.equal(out.cpu(), ref)
        assert torch.equal(hs.scale_coords(img1, coords.clone().to(cuda_device), img0, rp, round_=True).cpu()[:, :4], ref[:, :4].round())
    b = coords[:, :4].clone()
    port.clip_coords(b, (300, 500))
    o = coords[:, :4].clone().to(cuda_device)
    hs.clip_coords(o, (300, 500))
    assert torch.equal(o.cpu(), b)
    r = {'boxes': coords[:, :4].clone().to(cuda_device)}
    assert hs.rescale_outputs(r, 4.0) is r
    assert torch.equal(r['boxes'].cpu(), port.rescale_outputs({'boxes': coords[:, :4].clone()}, 4.0)['boxes'])
    # a column view of a wider tensor is modified in place (callers rely on it)
    wide = coords.clone().to(cuda_device)
    hs.clip_coords(wide[:, :4], (300, 500))
    assert torch.equal(wide[:, :4].cpu(), b) and torch.equal(wide[:, 4:].cpu(), coords[:, 4:])


@pytest.mark.parametrize("n_small,n_large,seed", [(20000, 60, 0), (5000, 400, 1)])
def test_merge_nms_with_large_boxes(cuda_device, n_small, n_large, seed):
    """Slide-level merge with boxes far wider than a hash cell (the warp-per-entry kernel of csrc/merge.cu) vs
    torchvision.ops.nms on the same rows (Ensemble.merge, yolo.py:189-195)."""
    import torchvision
    g = torch.Generator().manual_seed(seed)
    c = torch.rand((n_small, 2), generator=g) * 4000
    s = 12 + 24 * torch.rand((n_small, 2), generator=g)
    small = torch.cat([c - s / 2, c + s / 2], 1)
    centers = torch.rand((max(n_large // 6, 1), 2), generator=g) * 3000 + 500
    which = torch.randint(0, len(centers), (n_large,), generator=g)
    lc = centers[which] + 10 * torch.rand((n_large, 2), generator=g)
    ls = 100 + 900 * torch.rand((n_large, 1), generator=g) + 20 * torch.rand((n_large, 2), generator=g)
    large = torch.cat([lc - ls / 2, lc + ls / 2], 1)
    boxes = torch.cat([small, large])
    scores = torch.rand((len(boxes),), generator=g)
    perm = torch.randperm(len(boxes), generator=g)
    boxes, scores = boxes[perm].contiguous(), scores[perm].contiguous()
    conf, iou = 0.1, 0.45
    keep = scores > conf
    idx = torch.nonzero(keep).flatten()
    kept = idx[torchvision.ops.nms(boxes[idx], scores[idx], iou)]
    ref = torch.full((len(scores),), 2, dtype=torch.uint8)
    ref[~keep] = 3
    ref[kept] = 1
    got = hdy.merge_nms(boxes.to(cuda_device), scores.to(cuda_device), conf, iou).cpu()
    assert torch.equal(got, ref)


def _near_threshold_tile(g, n_pairs, n_free, thr, tile=1024.0, n_big=0):
    """Boxes of one tile (tile coordinates): pairs of equal squares shifted so that their IoU sits within ~2e-4 of thr
    (either side), plus free nuclei-sized boxes."""
    s = 14.0 + 16.0 * torch.rand((n_pairs, 1), generator=g)
    dx = s * (1 - thr) / (1 + thr) + (torch.rand((n_pairs, 1), generator=g) - 0.5) * 0.006
    c = 40.0 + (tile - 120.0) * torch.rand((n_pairs, 2), generator=g)
    a = torch.cat([c, c + s], 1)
    b = a + torch.cat([dx, torch.zeros_like(dx), dx, torch.zeros_like(dx)], 1)
    cf = tile * torch.rand((n_free, 2), generator=g)
    sf = 12.0 + 24.0 * torch.rand((n_free, 2), generator=g)
    free = torch.cat([cf - sf / 2, cf + sf / 2], 1)
    # big false positives hugging the tile border: they stick out by 0..200 px (adaptive far cap, far list, dirty
    # tiles) and land in the merge's large bucket (large -> small push)
    cb = tile * torch.rand((n_big, 2), generator=g)
    cb[:, 0] = torch.where(torch.rand((n_big,), generator=g) < 0.5, cb[:, 0] * 0.05, tile - cb[:, 0] * 0.05)
    sb = 50.0 + 350.0 * torch.rand((n_big, 2), generator=g)
    big = torch.cat([cb - sb / 2, cb + sb / 2], 1)
    boxes = torch.cat([a, b, free, big]).float()
    return boxes[torch.randperm(len(boxes), generator=g)].contiguous()


@pytest.mark.parametrize("thr,seed,n_big", [(0.45, 0, 0), (0.3, 1, 0), (0.7, 2, 0), (0.45, 3, 7), (0.3, 4, 3)])
def test_interior_shortcut_is_exact_far_from_the_origin(cuda_device, thr, seed, n_big):
    """At slide coordinates ~10^5 the fp32 rounding of `box + tile origin` (yolo_head.py:455) moves IoUs by ~5e-4, so
    Ensemble.merge (yolo.py:195) suppresses pairs the per-tile NMS kept.  The gray-zone flags of hdy_nms_tiles must
    route exactly those rows around the interior shortcut: verdicts == torchvision's dense NMS on the shifted boxes."""
    import torchvision
    from hd_yolo_b200 import ops, _lib
    from hd_yolo_b200.ops import DetectBatch
    dev = cuda_device
    g = torch.Generator().manual_seed(seed)
    rois_all = hs.sliding_window_scanner((100000, 100000), (1024, 1024), 64)
    n_cols = 105
    tiles = [r * n_cols + c for r in (100, 101, 102) for c in (100, 101, 102)]      # full tiles near (96000, 96000)
    rois = rois_all[tiles]
    bs, conf = len(tiles), 0.2
    per_tile = [_near_threshold_tile(g, 150, 120, thr, n_big=n_big) for _ in range(bs)]
    n = max(len(b) for b in per_tile)
    eps = float(np.spacing(np.float32(101000.0))) / 2
    cand = ops._Cand(dev, bs, n, tag="gz")
    boxes_pad = torch.zeros((bs, n, 4))
    scores_pad = torch.zeros((bs, n))
    for i, b in enumerate(per_tile):
        boxes_pad[i, :len(b)] = b
        scores_pad[i, :len(b)] = 0.25 + 0.7 * torch.rand((len(b),), generator=g)
    keys = torch.empty((bs, n), dtype=torch.int64, device=dev)
    ops._call("hdy_make_keys", _lib.ptr(scores_pad.to(dev)), bs, n, _lib.ptr(keys), ops._stream())
    cand.keys, cand.boxes = keys, boxes_pad.to(dev).contiguous()
    cand.counts = torch.tensor([len(b) for b in per_tile] + [0], dtype=torch.int32).to(dev)
    frag = []
    keep_idx, _, keep_box, keep_score, _, keep_counts, md = ops._run_nms(cand, thr, n, gray_eps=eps, fragile_out=frag)
    kc = keep_counts.cpu().tolist()
    # per-tile NMS == torchvision in tile coordinates
    kept_tiles, flips = [], 0
    for i, b in enumerate(per_tile):
        ref = torchvision.ops.nms(b, scores_pad[i, :len(b)], thr)
        assert torch.equal(keep_idx[i, :kc[i]].cpu().long(), ref)
        kb, ks = b[ref], scores_pad[i, :len(b)][ref]
        shifted = kb + torch.tensor([rois[i, 0], rois[i, 1], rois[i, 0], rois[i, 1]])
        flips += len(kb) - len(torchvision.ops.nms(shifted, ks, thr))
        kept_tiles.append((shifted, ks))
    assert flips > 0, "the construction must contain pairs that only the slide-level NMS suppresses"
    allb = torch.cat([t[0] for t in kept_tiles])
    alls = torch.cat([t[1] for t in kept_tiles])
    keep = alls > conf
    idx = torch.nonzero(keep).flatten()
    kept = idx[torchvision.ops.nms(allb[idx], alls[idx], thr)]
    ref_state = torch.full((len(alls),), 2, dtype=torch.uint8)
    ref_state[~keep] = 3
    ref_state[kept] = 1
    batch = DetectBatch(keep_box, None, keep_score, torch.zeros((bs, md), dtype=torch.int64, device=dev), None, None,
                        keep_idx, keep_counts, cand.counts, md, frag[0], eps, thr)
    acc = hs.SlideAccumulator(int(sum(kc)), dev)
    acc.append(batch, rois.to(dev))
    n_rows = acc.count()
    assert n_rows == len(alls)
    assert torch.equal(acc.boxes[:n_rows].cpu(), allb)
    fast = acc.verdicts(conf, thr, interior_shortcut=True)[:n_rows].cpu()
    full = acc.verdicts(conf, thr, interior_shortcut=False)[:n_rows].cpu()
    assert torch.equal(full, ref_state)
    assert torch.equal(fast, ref_state)
    # the shortcut really skipped most rows: fragile rows are a small minority
    n_frag = int(sum(int(frag[0][i, :kc[i]].sum()) for i in range(bs)))
    assert 0 < n_frag < 0.8 * n_rows


def test_scale_coords_vs_reference_golden(cuda_device):
    """C1 against what the reference's own scale_coords / clip_coords returned (tests/golden/scale_coords.npz)."""
    g = load_golden("scale_coords")
    for i in range(int(g["n"])):
        rp = g[f"rp{i}"].tolist()
        ratio_pad = ((rp[0], rp[0]), (rp[1], rp[2])) if rp[0] else None
        c = torch.from_numpy(g[f"in{i}"]).to(cuda_device)
        out = hs.scale_coords(tuple(g[f"img1_{i}"].tolist()), c, tuple(g[f"img0_{i}"].tolist()), ratio_pad)
        assert out.data_ptr() == c.data_ptr()                       # in place, like the reference
        assert torch.equal(out.cpu(), torch.from_numpy(g[f"out{i}"]))


def test_interior_shortcut_preconditions_are_enforced(cuda_device):
    """The shortcut is exact only under the conditions its gray-zone flags were produced for: a merge threshold below
    the per-tile NMS threshold, flags for a smaller coordinate range than the slide's, or a re-used accumulator whose
    tiling changed while the tile counts stayed the same must not silently keep interior rows."""
    from hd_yolo_b200 import synth
    dev = cuda_device
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    tile = 320

    def fill(acc, rois, eps, seed=3):
        dets = synth.slide_tile_logits(rois, tile, 4, seed=seed, device=dev, pitch=20.0)
        out = hdy.detect_postprocess(dets, spec, 0.25, 0.45, 600, gray_eps=eps)
        acc.append(out, rois.to(dev), rois_host=rois)
        return out

    rois_a = hs.sliding_window_scanner((600, 900), (tile, tile), 40)
    acc = hs.SlideAccumulator(6000, dev)
    eps = float(np.spacing(np.float32(1000.0))) / 2
    fill(acc, rois_a, eps)
    n = acc.count()
    fast = acc.verdicts(0.25, 0.45, interior_shortcut=True)[:n].clone()
    full = acc.verdicts(0.25, 0.45, interior_shortcut=False)[:n].clone()
    assert torch.equal(fast, full) and int((full == 2).sum()) > 0
    with pytest.raises(hdy.HdyError, match="below the per-tile NMS threshold"):
        acc.verdicts(0.25, 0.3, interior_shortcut=True)            # same-tile survivors may overlap more than 0.3
    # the same accumulator, a different tiling with the same number of batches and tiles: the cores are recomputed
    rois_b = rois_a.clone()
    rois_b[:, [0, 2]] += 5000.0
    acc.reset()
    fill(acc, rois_b, eps)
    n = acc.count()
    with pytest.raises(hdy.HdyError, match="round by up to"):       # flags made for coordinates up to ~1000 px
        acc.verdicts(0.25, 0.45, interior_shortcut=True)
    acc.reset()
    fill(acc, rois_b, float(np.spacing(np.float32(8000.0))) / 2)
    n = acc.count()
    fast = acc.verdicts(0.25, 0.45, interior_shortcut=True)[:n].clone()
    full = acc.verdicts(0.25, 0.45, interior_shortcut=False)[:n].clone()
    assert torch.equal(fast, full)


def test_captured_step_pins_its_scratch(cuda_device):
    """A CUDA graph bakes the addresses of its scratch buffers in: growing one of them under a live graph must raise,
    never free the memory the graph still writes to; eager work on another stream gets buffers of its own."""
    from hd_yolo_b200 import synth
    from hd_yolo_b200.ops import scratch_slot
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    small = [d.to(cuda_device) for d in synth.nuclei_logits(2, 160, 4, 100, seed=1, conf=0.25)]
    big = [d.to(cuda_device) for d in synth.nuclei_logits(8, 320, 4, 400, seed=2, conf=0.25)]
    step = hdy.CapturedStep(lambda: hdy.detect_postprocess(small, spec, 0.25, 0.45, 300, cap=512), slot=77)
    ref = [a['boxes'].clone() for a in step().to_list()]
    # the same slot, eagerly, on the default stream: its scratch is keyed by stream, the graph's buffers stay put
    with scratch_slot(77):
        hdy.detect_postprocess(big, spec, 0.25, 0.45, 300).to_list()
    again = [a['boxes'] for a in step().to_list()]
    assert all(torch.equal(a, b) for a, b in zip(ref, again))
    # on the graph's own stream a larger call would have to regrow the pinned buffers: refused
    with scratch_slot(77), torch.cuda.stream(step.stream):
        with pytest.raises(hdy.HdyError, match="CUDA graph"):
            hdy.detect_postprocess(big, spec, 0.25, 0.45, 300)
    torch.cuda.synchronize()


def test_tensors_of_another_device_are_refused(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    b = torch.zeros((4, 4), device="cuda:1")
    with pytest.raises(hdy.HdyError, match="current device"):
        hdy.nms(b, torch.zeros((4,), device="cuda:1"), 0.5)
