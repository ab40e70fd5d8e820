"""CPU tests: the oracle (oracle/port.py, oracle/nms_core.c) against the golden vectors generated from
the real reference (oracle/make_golden.py) and against known-answer cases."""
import numpy as np
import pytest
import torch
import torchvision

from conftest import load_golden, unpack_list
from oracle import nms_c, port
from hd_yolo_b200 import synth


def _dets(g):
    i, dets, preds = 0, [], []
    while f"det{i}" in g:
        dets.append(torch.from_numpy(g[f"det{i}"]))
        preds.append(torch.from_numpy(g[f"pred{i}"]))
        i += 1
    return dets, preds


@pytest.mark.parametrize("name", ["detect_640_l3", "detect_320_l4"])
def test_decode_matches_reference(name):
    g = load_golden(name)
    dets, preds = _dets(g)
    mine = port.compute_proposals(dets, g["anchors"].tolist(), g["strides"].tolist())
    for a, b in zip(mine, preds):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["detect_640_l3", "detect_320_l4"])
def test_compute_outputs_matches_reference(name):
    g = load_golden(name)
    _, preds = _dets(g)
    params = {'conf_thres': float(g["conf_thres"]), 'iou_thres': float(g["iou_thres"]), 'max_det': int(g["max_det"])}
    mine = port.compute_outputs([p.clone() for p in preds], int(g["nc"]), params)
    ref = unpack_list(g, "out", ["boxes", "scores", "labels"])
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert len(b["boxes"]) > 0
        for k in ("boxes", "scores", "labels"):
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_nms_per_image_matches_reference(tag):
    g = load_golden("nms_rows")
    preds = torch.from_numpy(g["preds"])
    conf, iou, md = g[f"npi_{tag}_params"].tolist()
    mine = port.nms_per_image(preds.clone(), int(g["nc"]), conf, iou, int(md))
    ref = unpack_list(g, f"npi_{tag}", ["boxes", "scores", "extra"])
    for a, b in zip(mine, ref):
        for k in ("boxes", "scores", "extra"):
            assert torch.equal(a[k], b[k])
    assert len(ref[2]["boxes"]) == 0  # the empty image stays empty


YOLO_KW = {"a": dict(conf_thres=0.25, iou_thres=0.45, max_det=300),
           "b": dict(conf_thres=0.1, iou_thres=0.45, multi_label=True, max_det=1000),
           "c": dict(conf_thres=0.25, iou_thres=0.5, agnostic=True, max_det=300),
           "d": dict(conf_thres=0.2, iou_thres=0.45, classes=[1, 3], max_det=300)}


@pytest.mark.parametrize("tag", list(YOLO_KW))
def test_non_max_suppression_matches_reference(tag):
    g = load_golden("nms_rows")
    nc = int(g["nc"])
    pred = torch.from_numpy(g["preds"])[..., :5 + nc].contiguous()
    mine = port.non_max_suppression(pred.clone(), **YOLO_KW[tag])
    sizes = g[f"yolo_{tag}_sizes"].tolist()
    assert [len(m) for m in mine] == sizes
    assert torch.equal(torch.cat(mine), torch.from_numpy(g[f"yolo_{tag}_out"]))


def test_hierarchical_scores_tree():
    g = load_golden("hier_tree")
    desc = {}
    for k, v in zip(g["ops_src"].tolist(), g["ops_dst"].tolist()):
        desc.setdefault(k, []).append(v)
    x = torch.from_numpy(g["scores_in"]).clone()
    assert torch.equal(port.hierarchical_scores(x, desc), torch.from_numpy(g["scores_out"]))


def test_tile_merge_matches_reference():
    g = load_golden("tile_merge")
    H, W = g["image_size"].tolist()
    rois = port.sliding_window_scanner((H, W), tuple(g["roi_size"].tolist()), int(g["overlap"]))
    assert torch.equal(rois, torch.from_numpy(g["rois"]))
    tiles = unpack_list(g, "tile", ["boxes", "scores", "labels"])
    for t, r in zip(tiles, rois):
        t["roi"] = r
    merged = port.merge_outputs(tiles)
    for k in ("boxes", "scores", "labels"):
        assert torch.equal(merged[k], torch.from_numpy(g["merged_" + k]))
    conf, iou, md = g["params"].tolist()
    final = port.ensemble_merge([{'det': merged}], {'conf_thres': conf, 'iou_thres': iou, 'max_det': md})['det']
    for k in ("boxes", "scores", "labels"):
        assert torch.equal(final[k], torch.from_numpy(g["final_" + k]))
    assert len(final["boxes"]) < len(merged["boxes"])  # the overlap bands did hold duplicates


def test_sliding_window_scanner_cases():
    g = load_golden("tile_merge")
    big = port.sliding_window_scanner((100000, 100000), (1024, 1024), 64)
    assert len(big) == int(g["scan_slide_n"]) == 11025
    assert torch.equal(big[:3], torch.from_numpy(g["scan_slide_head"]))
    assert torch.equal(big[-3:], torch.from_numpy(g["scan_slide_tail"]))
    assert big[-1].tolist() == [99840.0, 99840.0, 100000.0, 100000.0]
    assert torch.equal(port.sliding_window_scanner((300, 500), (128, 200), 0), torch.from_numpy(g["scan_small"]))
    assert torch.equal(port.sliding_window_scanner((100, 100), (128, 128), 16), torch.from_numpy(g["scan_fit"]))


def test_paste_masks_matches_reference():
    g = load_golden("paste_masks")
    H, W = g["shape"].tolist()
    out = port.paste_masks_in_image(torch.from_numpy(g["masks"]), torch.from_numpy(g["boxes"]), (H, W), padding=1)
    assert torch.equal(out, torch.from_numpy(g["out"]))


# ---------------------------------------------------------------- known answers for the third-party NMS
def _both(boxes, scores, thr):
    b = torch.tensor(boxes, dtype=torch.float32)
    s = torch.tensor(scores, dtype=torch.float32)
    tv = torchvision.ops.nms(b, s, thr).tolist()
    c = nms_c.nms(b.numpy(), s.numpy(), thr).tolist()
    assert tv == c
    return tv


def test_nms_known_answers():
    assert _both([[0, 0, 10, 10]] * 3, [.5, .5, .5], 0.5) == [0]                     # ties -> lowest index
    assert _both([[0, 0, 10, 10], [0, 0, 10, 5]], [.9, .8], 0.5) == [0, 1]            # iou == thr is kept
    assert _both([[0, 0, 10, 10], [0, 0, 10, 5]], [.9, .8], 0.4999) == [0]
    assert _both([[5, 5, 5, 5], [5, 5, 5, 5]], [.9, .8], 0.1) == [0, 1]               # 0/0 = NaN never suppresses
    assert _both([[0, 0, 4, 4], [10, 10, 14, 14], [0, 0, 4, 4]], [.1, .9, .5], 0.5) == [1, 2]  # score order


def test_nms_c_vs_torchvision_random():
    g = torch.Generator().manual_seed(11)
    for n, span, thr in [(1, 50, 0.5), (300, 60, 0.45), (2500, 320, 0.45), (2500, 200, 0.3), (800, 40, 0.6)]:
        c = torch.rand((n, 2), generator=g) * span
        wh = torch.rand((n, 2), generator=g) * 30 + 2
        b = torch.cat([c - wh / 2, c + wh / 2], 1)
        s = (torch.rand(n, generator=g) * 50).round() / 50
        assert torchvision.ops.nms(b, s, thr).tolist() == nms_c.nms(b.numpy(), s.numpy(), thr).tolist()


def test_conf_threshold_is_fp32():
    # `tensor_fp32 > 0.15` compares against float32(0.15)
    x = torch.tensor([np.float32(0.15)])
    assert not bool((x > 0.15).item())
    assert float(np.float32(0.15)) > 0.15


def test_synth_candidate_count():
    dets = synth.nuclei_logits(2, 320, 4, 250, seed=7, conf=0.25)
    preds = port.compute_proposals(dets, synth.ANCHORS_3, synth.STRIDES_3)
    cat = port.concat_levels(preds)
    n = (cat[..., 4] > 0.25).sum(1)
    assert ((n > 170) & (n < 330)).all()
    boxes = cat[..., 2:4][cat[..., 4] > 0.25]
    assert boxes.min() > 9 and boxes.max() < 40


def test_flatten_onehot_objects_matches_reference_formulation():
    """val_nuclei.py:34-48 restated with tile/repeat_interleave (the reference's formulation) vs the mirror."""
    import torch
    from hd_yolo_b200.ops import flatten_onehot_objects
    g = torch.Generator().manual_seed(3)
    k, nc1 = 37, 5
    x = {'labels': torch.rand((k, nc1), generator=g) > 0.6, 'boxes': torch.rand((k, 4), generator=g),
         'scores': torch.rand((k, nc1), generator=g), 'masks': torch.rand((k, 1, 7, 7), generator=g)}
    keep = x['labels'].flatten() > 0.
    ref_labels = torch.tile(torch.arange(nc1), (k,))[keep]
    ref_labels[ref_labels == 0] = -100
    got = flatten_onehot_objects(x)
    assert torch.equal(got['labels'], ref_labels)
    assert torch.equal(got['boxes'], torch.repeat_interleave(x['boxes'], nc1, 0)[keep])
    assert torch.equal(got['scores'], x['scores'].flatten()[keep])
    assert torch.equal(got['masks'], torch.repeat_interleave(x['masks'], nc1, 0)[keep])


# ------------------------------------------------------------------------ next rows: RoIAlign, AP matching
def _roi_golden():
    g = load_golden("roi_align")
    feats = [torch.from_numpy(g[f"feat{i}"]) for i in range(3)]
    return g, feats, torch.from_numpy(g["boxes"]), torch.from_numpy(g["levels"])


def test_multiscale_roi_align_matches_reference():
    g, feats, boxes, levels = _roi_golden()
    mine = port.multiscale_roi_align(feats, boxes, levels, g["strides"].tolist())
    assert torch.equal(mine, torch.from_numpy(g["out"]))
    assert float(mine[5].abs().max()) == 0.0 and float(mine[2].abs().max()) == 0.0   # no such level / fully outside


def test_roi_align_c_restatement_is_bit_exact():
    """oracle/roi_align_core.c against the golden (reference output) and against torchvision on wider cases."""
    from oracle import roi_align_c
    g, feats, boxes, levels = _roi_golden()
    ref = g["out"]
    for i, s in enumerate(g["strides"].tolist()):
        idx = np.where(g["levels"] == i)[0]
        mine = roi_align_c.roi_align(g[f"feat{i}"], g["boxes"][idx], 14, 1.0 / s)
        assert np.array_equal(mine, ref[idx])
    gen = torch.Generator().manual_seed(11)
    f = torch.randn((2, 3, 9, 13), generator=gen)
    c = torch.rand((64, 2), generator=gen) * torch.tensor([13 * 8., 9 * 8.])
    s = torch.rand((64, 2), generator=gen) * 70 + 1
    rois = torch.cat([torch.randint(0, 2, (64, 1), generator=gen).float(), c - s / 2, c + s / 2], 1)
    for M, S, aligned in [(14, 2, False), (7, 2, True), (14, 1, False), (5, 3, False), (16, 4, True)]:
        tv = torchvision.ops.roi_align(f, rois, (M, M), 1 / 8, S, aligned)
        mine = roi_align_c.roi_align(f.numpy(), rois.numpy(), M, 1 / 8, S, aligned)
        assert np.array_equal(mine, tv.numpy()), (M, S, aligned)


def test_apmeter_matches_reference():
    g = load_golden("ap_match")
    st = port.APMeterState()
    for i in range(int(g["n_images"])):
        out = {k: torch.from_numpy(g[f"out{i}_{k}"]) for k in ("boxes", "scores", "labels")}
        tgt = {k: torch.from_numpy(g[f"tgt{i}_{k}"]) for k in ("boxes", "labels")}
        st.add(out, tgt)
    assert [st.n_pred, st.n_true, st.n_match] == g["meter_n"].tolist()
    assert st.n_match > 40
    for f in ("scores", "y_pred", "y_true", "ious", "m_pred", "m_true"):
        assert torch.equal(getattr(st, f), torch.from_numpy(g["meter_" + f])), f


def test_box_iou_known_answers():
    a = torch.tensor([[0., 0., 10., 10.], [5., 5., 5., 5.]])
    b = torch.tensor([[0., 0., 10., 5.], [20., 20., 30., 30.], [5., 5., 5., 5.]])
    iou = port.box_iou(a, b)
    assert iou[0, 0] == 0.5 and iou[0, 1] == 0.0
    assert torch.isnan(iou[1, 2])                  # 0 / 0: never `>= 0.5`


# ------------------------------------------------------------------------------------------ scale_coords golden
def _scale_cases():
    g = load_golden("scale_coords")
    for i in range(int(g["n"])):
        rp = g[f"rp{i}"].tolist()
        ratio_pad = ((rp[0], rp[0]), (rp[1], rp[2])) if rp[0] else None
        yield (tuple(g[f"img1_{i}"].tolist()), torch.from_numpy(g[f"in{i}"]), tuple(g[f"img0_{i}"].tolist()), ratio_pad,
               torch.from_numpy(g[f"out{i}"]))


def test_scale_coords_port_vs_reference_golden():
    """C1 (utils_general.py:161-190): the port gives what the reference itself returned (bit for bit)."""
    n = 0
    for img1, c, img0, rp, want in _scale_cases():
        assert torch.equal(port.scale_coords(img1, c.clone(), img0, rp), want)
        n += 1
    assert n == 4


# ------------------------------------------------------------------------------------------ process_mask known answers
def test_process_mask_known_answers():
    """process_mask / crop_mask are upstream ultralytics/yolov5 v7.0 (utils/segment/general.py), which the reference
    does not contain and this container cannot fetch (no network): the port is a restatement of the published
    algorithm, PARITY UNPINNED.  What can be pinned are hand-computable consequences of that algorithm:
      * sigmoid(coef . protos) > 0.5  <=>  coef . protos > 0: with one-hot coefficients the mask is `proto_c > 0`;
      * crop_mask keeps x1 <= col < x2, y1 <= row < y2 on the box scaled by (mw / iw, mh / ih), fractional edges
        included exactly as float comparisons against arange (a box edge at 2.5 keeps column 3, not column 2);
      * upsample: bilinear, align_corners=False, thresholded AFTER the interpolation: a single proto pixel of value 1
        (sigmoid(+large)) among zeros stays > 0.5 only where its weight exceeds 0.5."""
    mh = mw = 8
    protos = torch.full((2, mh, mw), -20.0)
    protos[0, 2:6, 1:7] = 20.0                  # channel 0: a 4 x 6 blob
    protos[1, 4, 4] = 20.0                      # channel 1: one pixel
    coef = torch.tensor([[1.0, 0.0], [0.0, 1.0]])
    full = torch.tensor([[0.0, 0.0, 32.0, 32.0], [0.0, 0.0, 32.0, 32.0]])
    m = port.process_mask(protos, coef, full.clone(), (32, 32), upsample=False)
    assert torch.equal(m[0], (protos[0] > 0).float()) and torch.equal(m[1], (protos[1] > 0).float())
    # crop: box [10, 6, 22, 18] px at 1/4 scale = [2.5, 1.5, 5.5, 4.5] -> columns 3..5, rows 2..4
    box = torch.tensor([[10.0, 6.0, 22.0, 18.0]])
    c = port.process_mask(protos, coef[:1], box.clone(), (32, 32), upsample=False)[0]
    want = torch.zeros((mh, mw))
    want[2:5, 3:6] = 1.0
    assert torch.equal(c, want)
    # upsample of the single pixel (channel 1, proto (4, 4)): at 4x, output pixel o has src = 0.25 * (o + 0.5) - 0.5; the
    # weight of proto pixel 4 is 1 - |src - 4| -> > 0.5 for o in 17..18 and exactly 0.5 (not kept) at o = 15.5/19.5: none
    u = port.process_mask(protos, coef[1:], full[1:].clone(), (32, 32), upsample=True)[0]
    rows = torch.nonzero(u.sum(1)).flatten().tolist()
    cols = torch.nonzero(u.sum(0)).flatten().tolist()
    # weight_y * weight_x > 0.5 (sigmoid(20) ~ 1, sigmoid(-20) ~ 0): per-axis weights are 0.625, 0.875, 0.875, 0.625
    # for o = 16..19, so the product exceeds 0.5 for the pairs with both weights 0.875 and the mixed 0.875 * 0.625 ones
    w = {16: 0.625, 17: 0.875, 18: 0.875, 19: 0.625}
    expect = torch.zeros((32, 32))
    for y, wy in w.items():
        for x, wx in w.items():
            if wy * wx > 0.5:
                expect[y, x] = 1.0
    assert torch.equal(u, expect) and rows == [16, 17, 18, 19] and cols == [16, 17, 18, 19]
