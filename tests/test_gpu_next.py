"""GPU parity tests of the rows SURVEY 8f ranks next: multi-scale RoIAlign (the gather in front of the reference's
mask head) and the prediction/ground-truth matching of APMeter.add.  Oracle: oracle/port.py + oracle/roi_align_core.c,
pinned on reference-generated goldens (tests/golden/roi_align.npz, ap_match.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import hd_yolo_b200 as hdy
from hd_yolo_b200 import synth
from oracle import port, roi_align_c

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ RoIAlign
def test_roi_align_golden_bit_exact(cuda_device):
    g = load_golden("roi_align")
    feats = [torch.from_numpy(g[f"feat{i}"]).to(cuda_device) for i in range(3)]
    boxes, levels = torch.from_numpy(g["boxes"]).to(cuda_device), torch.from_numpy(g["levels"]).to(cuda_device)
    out = hdy.multiscale_roi_align(feats, boxes, levels, g["strides"].tolist()).cpu()
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    assert torch.equal(out, ref), f"max abs diff {(out - ref).abs().max().item()}"
    assert float(out[5].abs().max()) == 0.0          # level id 3: the zero row of the reference's `result`


@pytest.mark.parametrize("C,M,S,aligned,lo,hi", [
    (256, 14, 2, False, 12, 36),    # the reference's call on nuclei: all on the tensor-core path but a few 7-tap windows
    (128, 14, 2, False, 8, 120),    # mixed: the large RoIs go to the exact kernel
    (128, 7, 2, True, 4, 40),       # one M tile (49 bins)
    (256, 16, 3, False, 6, 30),     # 256 bins: over the 200 rows of the A operand, the exact kernel does them all
    (128, 13, 3, False, 6, 30),     # S = 3 (weights are thirds)
    (128, 11, 1, True, 6, 30),
])
@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_align_tf32x3_within_tolerance(cuda_device, C, M, S, aligned, lo, hi, channels_last):
    """mode='tf32x3' (tcgen05 3xTF32, roi_align_tc.cu) against the exact-order kernel on the same inputs: the north
    star's 1e-5 relative tolerance, measured against the magnitude of the taps (|w| . |f|, which bounds every partial
    sum); rows the reference leaves zero stay exactly zero; RoIs over 6 x 6 taps are bit-identical (exact kernel)."""
    gen = torch.Generator().manual_seed(C + 7 * M + S)
    tile, bs, K = 320, 3, 700
    strides = [8, 16, 32]
    feats = [torch.randn((bs, C, tile // s, tile // s), generator=gen).to(cuda_device) for s in strides]
    rois = _rois(gen, K, bs, tile, lo, hi)
    rois[0, 1:] = torch.tensor([-40., -40., -20., -20.])           # outside: all weights zero
    rois[1, 1:] = torch.tensor([tile - 3., tile - 3., tile + 30., tile + 30.])
    rois[4, 1:] = torch.tensor([100., 100., 100., 100.])           # empty box (aligned=False clamps to 1 px)
    levels = torch.randint(0, 3, (K,), generator=gen).float()
    levels[2], levels[3] = -1.0, 0.5                                # match no level
    levels[0] = 0.0                                                 # (on a coarse level the far box is in range again)
    rois, levels = rois.to(cuda_device), levels.to(cuda_device)
    exact = hdy.multiscale_roi_align(feats, rois, levels, strides, M, S, aligned)
    feats_in = [f.contiguous(memory_format=torch.channels_last) for f in feats] if channels_last else feats
    fast = hdy.multiscale_roi_align(feats_in, rois, levels, strides, M, S, aligned, mode="tf32x3")
    absf = hdy.multiscale_roi_align([f.abs() for f in feats], rois, levels, strides, M, S, aligned)   # sum |w| |f|
    err = (fast - exact).abs()
    assert bool((err <= 1e-5 * absf + 1e-30).all()), \
        f"max err {err.max().item():.3g}, worst ratio {(err / absf.clamp_min(1e-30)).max().item():.3g}"
    assert float(fast[2].abs().max()) == 0.0 and float(fast[3].abs().max()) == 0.0 and float(fast[0].abs().max()) == 0.0
    same = (fast == exact).flatten(1).all(1)
    assert 0 < int(same.sum()) and (int(same.sum()) < K or M * M > 200)   # some rows took the exact kernel
    if M * M > 200:
        assert torch.equal(fast, exact)
    else:
        assert float(err.max()) > 0.0                 # ... so the tensor-core path did run


def test_roi_align_tf32x3_golden(cuda_device):
    g = load_golden("roi_align")
    feats = [torch.from_numpy(g[f"feat{i}"]).to(cuda_device) for i in range(3)]
    feats = [f.repeat(1, 64 // f.shape[1] + 1, 1, 1)[:, :64].contiguous() for f in feats]
    boxes, levels = torch.from_numpy(g["boxes"]).to(cuda_device), torch.from_numpy(g["levels"]).to(cuda_device)
    exact = hdy.multiscale_roi_align(feats, boxes, levels, g["strides"].tolist())
    fast = hdy.multiscale_roi_align(feats, boxes, levels, g["strides"].tolist(), mode="tf32x3")
    assert torch.allclose(fast, exact, rtol=1e-5, atol=1e-5)
    fast_cl = hdy.multiscale_roi_align([f.contiguous(memory_format=torch.channels_last) for f in feats], boxes, levels,
                                       g["strides"].tolist(), mode="tf32x3")
    assert torch.allclose(fast_cl, exact, rtol=1e-5, atol=1e-5)
    with pytest.raises(hdy.HdyError):
        hdy.multiscale_roi_align([f[:, :40].contiguous() for f in feats], boxes, levels, g["strides"].tolist(),
                                 mode="tf32x3")
    with pytest.raises(hdy.HdyError):       # the exact kernel reads the reference's layout only
        hdy.multiscale_roi_align([f.contiguous(memory_format=torch.channels_last) for f in feats], boxes, levels,
                                 g["strides"].tolist())


def _rois(gen, K, bs, tile, lo, hi):
    c = torch.rand((K, 2), generator=gen) * tile
    s = torch.rand((K, 2), generator=gen) * (hi - lo) + lo
    return torch.cat([torch.randint(0, bs, (K, 1), generator=gen).float(), c - s / 2, c + s / 2], 1)


@pytest.mark.parametrize("C,M,S,aligned,lo,hi", [
    (40, 14, 2, False, 8, 48),      # the reference's call; ragged channel chunk (32 + 8)
    (37, 14, 2, False, 8, 48),      # odd channel count: ragged pair
    (8, 14, 2, False, 100, 500),    # windows too large to stage: in-place path
    (33, 7, 2, True, 4, 64),
    (16, 14, 1, False, 4, 64),
    (6, 5, 3, False, 4, 200),
    (5, 16, 4, True, 2, 90),
])
def test_roi_align_matches_oracle(cuda_device, C, M, S, aligned, lo, hi):
    gen = torch.Generator().manual_seed(C * 100 + M)
    tile, bs, K = 320, 3, 300
    strides = [8, 16, 32]
    feats = [torch.randn((bs, C, tile // s, tile // s), generator=gen) for s in strides]
    rois = _rois(gen, K, bs, tile, lo, hi)
    rois[0, 1:] = torch.tensor([-40., -40., -20., -20.])           # outside
    rois[1, 1:] = torch.tensor([tile - 3., tile - 3., tile + 30., tile + 30.])
    levels = torch.randint(0, 3, (K,), generator=gen).float()
    levels[2], levels[3] = -1.0, 0.5                                # match no level
    out = hdy.multiscale_roi_align([f.to(cuda_device) for f in feats], rois.to(cuda_device), levels.to(cuda_device),
                                   strides, M, S, aligned).cpu().numpy()
    ref = np.zeros_like(out)
    for i, s in enumerate(strides):
        idx = np.where(levels.numpy() == i)[0]
        ref[idx] = roi_align_c.roi_align(feats[i].numpy(), rois.numpy()[idx], M, 1.0 / s, S, aligned)
    assert np.array_equal(out, ref), f"max abs diff {np.abs(out - ref).max()}"
    # and the oracle itself is torchvision's op (float tolerance as stated by north_star: 1e-5 relative)
    tv = port.multiscale_roi_align(feats, rois, levels, strides, M, S, aligned).numpy()
    assert np.allclose(out, tv, rtol=1e-5, atol=1e-6)


def test_roi_align_single_level_and_list_boxes(cuda_device):
    import torchvision
    gen = torch.Generator().manual_seed(5)
    f = torch.randn((2, 12, 40, 40), generator=gen)
    per_image = [_rois(gen, 17, 1, 320, 8, 60)[:, 1:], _rois(gen, 9, 1, 320, 8, 60)[:, 1:]]
    out = hdy.roi_align(f.to(cuda_device), [b.to(cuda_device) for b in per_image], (14, 14), 1 / 8, 2, False).cpu()
    ref = torchvision.ops.roi_align(f, per_image, (14, 14), 1 / 8, 2, False)
    assert torch.equal(out, ref)
    empty = hdy.roi_align(f.to(cuda_device), torch.zeros((0, 5), device=cuda_device), 14, 1 / 8)
    assert empty.shape == (0, 12, 14, 14)


def test_roi_align_after_detect_postprocess(cuda_device):
    """compute_outputs' mask front half (yolo_head.py:320-330): proposals + level ids of a DetectBatch -> features."""
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    dets = synth.nuclei_logits(2, 320, 4, 200, seed=9, conf=0.25)
    out = hdy.detect_postprocess([d.to(cuda_device) for d in dets], spec, 0.25, 0.45, 300)
    rois, lv = hdy.batch_rois(out)
    gen = torch.Generator().manual_seed(3)
    feats = [torch.randn((2, 16, 320 // s, 320 // s), generator=gen) for s in synth.STRIDES_3]
    got = hdy.multiscale_roi_align([f.to(cuda_device) for f in feats], rois, lv, synth.STRIDES_3).cpu()
    ref = port.multiscale_roi_align(feats, rois.cpu(), lv.cpu(), synth.STRIDES_3)
    assert len(rois) == int(out.counts.sum()) and len(rois) > 100
    assert torch.equal(got, ref)


# ------------------------------------------------------------------------------------------------ matching
def _meter_inputs(g, i):
    out = {k: torch.from_numpy(g[f"out{i}_{k}"]) for k in ("boxes", "scores", "labels")}
    tgt = {k: torch.from_numpy(g[f"tgt{i}_{k}"]) for k in ("boxes", "labels")}
    return out, tgt


def test_apmeter_golden(cuda_device):
    g = load_golden("ap_match")
    meter = hdy.APMeter()
    for i in range(int(g["n_images"])):
        out, tgt = _meter_inputs(g, i)
        meter.add({k: v.to(cuda_device) for k, v in out.items()}, {k: v.to(cuda_device) for k, v in tgt.items()})
    assert [meter.n_pred, meter.n_true, meter.n_match] == g["meter_n"].tolist()
    for f in ("scores", "y_pred", "y_true", "ious", "m_pred", "m_true"):
        assert torch.equal(getattr(meter, f), torch.from_numpy(g["meter_" + f])), f


def test_box_iou_matches_oracle(cuda_device):
    gen = torch.Generator().manual_seed(2)
    a = _rois(gen, 700, 1, 300, 5, 60)[:, 1:].contiguous()
    b = _rois(gen, 450, 1, 300, 5, 60)[:, 1:].contiguous()
    a[3] = torch.tensor([5., 5., 5., 5.])
    b[7] = torch.tensor([5., 5., 5., 5.])          # 0/0 -> NaN in both
    got = hdy.box_iou(a.to(cuda_device), b.to(cuda_device)).cpu()
    ref = port.box_iou(a, b)
    assert torch.equal(torch.isnan(got), torch.isnan(ref)) and bool(torch.isnan(got[3, 7]))
    assert torch.equal(torch.nan_to_num(got, nan=-1.0), torch.nan_to_num(ref, nan=-1.0))
    assert hdy.box_iou(a[:0].to(cuda_device), b.to(cuda_device)).shape == (0, 450)


@pytest.mark.parametrize("k,n_gt,seed", [(1500, 1400, 1), (5000, 300, 2), (40, 4500, 3)])
def test_match_predictions_matches_oracle(cuda_device, k, n_gt, seed):
    gen = torch.Generator().manual_seed(seed)
    gt = _rois(gen, n_gt, 1, 1024, 12, 36)[:, 1:].contiguous()
    m = min(k, n_gt)
    pb = torch.cat([gt[:m] + torch.randn((m, 4), generator=gen) * 1.5, _rois(gen, k - m, 1, 1024, 12, 36)[:, 1:]])
    pb = pb[torch.randperm(k, generator=gen)].contiguous()
    scores = (torch.rand(k, generator=gen) * 256).round() / 256          # score ties: lower row first
    out = {'boxes': pb, 'scores': scores, 'labels': torch.randint(-1, 5, (k,), generator=gen)}
    tgt = {'boxes': gt, 'labels': torch.randint(1, 5, (n_gt,), generator=gen)}
    got = hdy.match_predictions({a: b.to(cuda_device) for a, b in out.items()},
                                {a: b.to(cuda_device) for a, b in tgt.items()}, 0.5)
    s, l, p, t, i = port.apmeter_match(out, tgt, 0.5)
    assert len(i) > m // 4
    assert torch.equal(got['scores'].cpu(), s) and torch.equal(got['labels'].cpu(), l)
    assert torch.equal(got['ious'].cpu(), i)
    assert torch.equal(got['pred_idx'].cpu(), p) and torch.equal(got['true_idx'].cpu(), t)


def test_match_predictions_ties_and_capacity(cuda_device):
    # every prediction is an exact copy of every ground-truth box: k*g pairs of iou 1, more than the default capacity
    box = torch.tensor([[10., 10., 40., 40.]])
    k, n_gt = 70, 60
    out = {'boxes': box.repeat(k, 1), 'scores': torch.linspace(0.9, 0.1, k), 'labels': torch.ones(k, dtype=torch.int64)}
    tgt = {'boxes': box.repeat(n_gt, 1), 'labels': torch.ones(n_gt, dtype=torch.int64)}
    got = hdy.match_predictions({a: b.to(cuda_device) for a, b in out.items()},
                                {a: b.to(cuda_device) for a, b in tgt.items()}, 0.5, cap=64)
    assert len(got['ious']) == k * n_gt and bool((got['ious'] == 1.0).all())
    idx = torch.arange(k * n_gt)
    assert torch.equal(got['pred_idx'].cpu(), idx // n_gt) and torch.equal(got['true_idx'].cpu(), idx % n_gt)


def test_compute_outputs_with_masks_composed(cuda_device):
    """Detect.compute_outputs(compute_masks=True) (yolo_head.py:301-355) end to end: fused decode + nms_per_image, the
    proposals / levels pair, multiscale RoIAlign, the reference's own mask head (a PyTorch module, here two small
    convolutions), sigmoid + per-label channel select.  Checker: the oracle composition on the device-decoded rows
    (the reference's own :348 raises on torch >= 2; the port restates it with the integer clamp it means)."""
    import torch.nn as nn
    from hd_yolo_b200 import synth
    torch.manual_seed(11)
    torch.backends.cudnn.allow_tf32 = False
    dev = cuda_device
    nc, C = 4, 8
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=nc)
    dets = synth.nuclei_logits(2, 320, nc, 250, seed=4, conf=0.25)
    feats = [torch.randn((2, C, 320 // s, 320 // s)) for s in synth.STRIDES_3]
    seg_h = nn.Sequential(nn.Conv2d(C, C, 3, padding=1), nn.ReLU(), nn.ConvTranspose2d(C, C, 2, 2), nn.ReLU(),
                          nn.Conv2d(C, 2, 1)).eval()
    mask_indices = torch.tensor([-1, 0, 0, 1, 1])
    params = {'conf_thres': 0.25, 'iou_thres': 0.45, 'max_det': 300}
    with torch.no_grad():
        got = hdy.compute_outputs([d.to(dev) for d in dets], spec, [f.to(dev) for f in feats], True,
                                  seg_h=seg_h.to(dev), mask_indices=mask_indices.to(dev), nms_params=params)
        seg_h = seg_h.cpu()
        cat = hdy.decode_concat([d.to(dev) for d in dets], spec).cpu()
        outs = port.nms_per_image(cat, nc, 0.25, 0.45, 300)
        n_masks = 0
        for i, (g, o) in enumerate(zip(got, outs)):
            s, l = port.select_scores(o['scores'].clone(), 0.25, port.default_descendants(nc))
            assert torch.equal(g['boxes'].cpu(), o['boxes']) and torch.equal(g['labels'].cpu(), l)
            assert torch.equal(g['scores'].cpu(), s)
            k = len(o['boxes'])
            assert k > 50
            rois = torch.nn.functional.pad(o['boxes'], [1, 0], value=float(i))
            mf = port.multiscale_roi_align(feats, rois, o['extra'][:, 0], synth.STRIDES_3)
            ref = port.mask_select(seg_h(mf), l, mask_indices)
            assert g['masks'].shape == (k, 1, 28, 28)
            assert float((g['masks'].cpu() - ref).abs().max()) < 2e-5
            assert float(g['masks'].cpu()[l < 0].abs().sum()) == 0.0      # unclassified (-100 -> index -1): no mask
            n_masks += k
    assert n_masks > 100
    no_masks = hdy.compute_outputs([d.to(dev) for d in dets], spec, compute_masks=False, nms_params=params)
    assert all('masks' not in r for r in no_masks) and len(no_masks) == 2
