"""ORACLE (test infrastructure): generate tests/golden/*.npz from the REAL reference.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
Every fixture stores the seeded inputs and what the reference's own functions returned for them
(torch 2.11.0 CPU, torchvision 0.26.0, fp32).  The fixtures pin oracle/port.py (tests/test_oracle.py)
and are compared against the CUDA path in the -m gpu tests.  Nothing here is imported by the product.
"""
import os
import sys

import numpy as np
import torch

from . import ref_shim
from hd_yolo_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KB")


def _pack_list(prefix, lst, keys):
    """list of dicts -> flat arrays + offsets"""
    out = {}
    sizes = [len(d[keys[0]]) for d in lst]
    out[prefix + "_sizes"] = np.array(sizes, dtype=np.int64)
    for k in keys:
        out[prefix + "_" + k] = _np(torch.cat([d[k] for d in lst])) if lst else np.zeros((0,))
    return out


def make_detect(ref, name, tile, strides, anchors, nc, seed, conf, max_det):
    """cfg 1 (BASELINE.json configs[0]): random-init reference Detect head on randn features."""
    torch.manual_seed(seed)
    ch = [16, 24, 32, 40][:len(strides)]
    det = ref.Detect(ch=ch, anchors=anchors, strides=strides, nc=nc, masks={}, is_scripting=True)
    det.f = list(range(len(strides)))
    det.eval()
    det.nms_params = det.get_nms_params({'conf_thres': conf, 'max_det': max_det})
    feats = {i: torch.randn(1, c, tile // s, tile // s) for i, (c, s) in enumerate(zip(ch, strides))}
    with torch.no_grad():
        dets = []
        for i, m in enumerate(det.m):
            f = m(feats[i])
            bs, _, ny, nx = f.shape
            dets.append(f.view(bs, det.na, det.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous())
        preds = det.compute_proposals([d.clone() for d in dets])
        outs = det.compute_outputs([p.clone() for p in preds], [], compute_masks=False)
    arrays = {"tile": np.int64(tile), "nc": np.int64(nc), "strides": np.array(strides, dtype=np.float32),
              "anchors": np.array(anchors, dtype=np.float32), "conf_thres": np.float64(conf),
              "iou_thres": np.float64(det.nms_params['iou_thres']), "max_det": np.int64(max_det)}
    for i, (d, p) in enumerate(zip(dets, preds)):
        arrays[f"det{i}"] = _np(d)
        arrays[f"pred{i}"] = _np(p)
    arrays.update(_pack_list("out", outs, ["boxes", "scores", "labels"]))
    _save(name, **arrays)


def make_nms(ref, name, seed):
    """Synthetic decoded rows: nuclei-sized boxes, score ties, tiny boxes, one empty image."""
    g = torch.Generator().manual_seed(seed)
    bs, N, nc, E = 4, 1500, 4, 1
    c = torch.rand((bs, N, 2), generator=g) * 320
    wh = torch.rand((bs, N, 2), generator=g) * 30 + 1.0       # some sides < 2 px
    obj = torch.rand((bs, N, 1), generator=g)
    obj = (obj * 64).round() / 64                              # many exact ties
    cls = torch.rand((bs, N, nc), generator=g)
    extra = torch.randint(0, 3, (bs, N, E), generator=g).float()
    preds = torch.cat([c, wh, obj, cls, extra], -1)
    preds[2, :, 4] = 0.01                                      # image 2: nothing passes conf
    preds[3, 100:110] = preds[3, 100:101]                      # exact duplicate rows
    arrays = {"preds": _np(preds), "nc": np.int64(nc)}
    for tag, kw in {"a": dict(conf_thres=0.25, iou_thres=0.45, max_det=300),
                    "b": dict(conf_thres=0.5, iou_thres=0.3, max_det=10000),
                    "c": dict(conf_thres=0.05, iou_thres=0.6, max_det=50)}.items():
        outs = ref.nms_per_image(preds.clone(), nc=nc, **kw)
        arrays.update(_pack_list("npi_" + tag, outs, ["boxes", "scores", "extra"]))
        arrays["npi_" + tag + "_params"] = np.array([kw['conf_thres'], kw['iou_thres'], kw['max_det']])
    pred5 = preds[..., :5 + nc].contiguous()
    for tag, kw in {"a": dict(conf_thres=0.25, iou_thres=0.45, max_det=300),
                    "b": dict(conf_thres=0.1, iou_thres=0.45, multi_label=True, max_det=1000),
                    "c": dict(conf_thres=0.25, iou_thres=0.5, agnostic=True, max_det=300),
                    "d": dict(conf_thres=0.2, iou_thres=0.45, classes=[1, 3], max_det=300)}.items():
        outs = ref.non_max_suppression(pred5.clone(), **kw)
        arrays["yolo_" + tag + "_sizes"] = np.array([len(o) for o in outs], dtype=np.int64)
        arrays["yolo_" + tag + "_out"] = _np(torch.cat(outs))
    _save(name, **arrays)


def make_hier(ref, name, seed):
    """Non-default class tree: hierarchical_scores order of products (yolo_head.py:473-491)."""
    tree = {0: {1: {4: {}, 5: {}}, 2: {6: {}}, 3: {}}}

    class TreeDetect(ref.Detect):
        def build_hierarchical_tree(self):
            return tree

    torch.manual_seed(seed)
    nc = 6
    det = TreeDetect(ch=[8, 8, 8], anchors=synth.ANCHORS_3, strides=synth.STRIDES_3, nc=nc, masks={}, is_scripting=True)
    x = torch.rand(257, 1 + nc)
    y = det.hierarchical_scores(x.clone())
    ks, vs = [], []
    for k, v in det.descendants.items():
        for vv in v:
            ks.append(k)
            vs.append(vv)
    _save(name, scores_in=_np(x), scores_out=_np(y), ops_src=np.array(ks, dtype=np.int64),
          ops_dst=np.array(vs, dtype=np.int64))


def make_merge(ref, name, seed):
    """sliding_window_scanner + Detect.merge_outputs + Ensemble.merge on a small synthetic slide."""
    g = torch.Generator().manual_seed(seed)
    H = W = 700
    rois = ref.sliding_window_scanner((H, W), (256, 256), 64)
    # global nuclei field; every tile re-detects the nuclei it contains with <= 1 px jitter
    nuc = torch.rand((900, 2), generator=g) * 700
    size = torch.rand((900, 2), generator=g) * 24 + 12
    tiles = []
    for roi in rois:
        x0, y0, x1, y1 = roi.tolist()
        inside = (nuc[:, 0] > x0 + 4) & (nuc[:, 0] < x1 - 4) & (nuc[:, 1] > y0 + 4) & (nuc[:, 1] < y1 - 4)
        c = nuc[inside] - torch.tensor([x0, y0]) + (torch.rand((int(inside.sum()), 2), generator=g) - 0.5)
        s = size[inside]
        boxes = torch.cat([c - s / 2, c + s / 2], 1)
        scores = torch.rand(len(boxes), generator=g) * 0.8 + 0.1
        scores = (scores * 128).round() / 128
        labels = torch.randint(1, 5, (len(boxes),), generator=g)
        tiles.append({'boxes': boxes, 'scores': scores, 'labels': labels, 'roi': roi})
    merged = ref.Detect.merge_outputs(None, [dict(t) for t in tiles])
    ens = ref.Ensemble([], nms_params={'conf_thres': 0.2, 'iou_thres': 0.45, 'max_det': 100000})
    final = ens.merge([{'det': merged}])['det']
    arrays = {"rois": _np(rois), "image_size": np.array([H, W]), "roi_size": np.array([256, 256]),
              "overlap": np.int64(64), "params": np.array([0.2, 0.45, 100000])}
    arrays.update(_pack_list("tile", tiles, ["boxes", "scores", "labels"]))
    for k in ("boxes", "scores", "labels"):
        arrays["merged_" + k] = _np(merged[k])
        arrays["final_" + k] = _np(final[k])
    # scanner-only cases (incl. the slide config: 11 025 tiles, last row [99840, 99840, 1e5, 1e5])
    big = ref.sliding_window_scanner((100000, 100000), (1024, 1024), 64)
    arrays["scan_slide_n"] = np.int64(len(big))
    arrays["scan_slide_head"] = _np(big[:3])
    arrays["scan_slide_tail"] = _np(big[-3:])
    arrays["scan_small"] = _np(ref.sliding_window_scanner((300, 500), (128, 200), 0))
    arrays["scan_fit"] = _np(ref.sliding_window_scanner((100, 100), (128, 128), 16))
    _save(name, **arrays)


def make_paste(ref, name, seed):
    """paste_masks_in_image as called at val_nuclei.py:169-176 / evaluation.py:122-123."""
    g = torch.Generator().manual_seed(seed)
    k, H, W = 14, 96, 120
    masks = torch.rand((k, 1, 28, 28), generator=g)
    c = torch.rand((k, 2), generator=g) * torch.tensor([W, H])
    s = torch.rand((k, 2), generator=g) * 40 + 3
    boxes = torch.cat([c - s / 2, c + s / 2], 1)
    boxes[0] = torch.tensor([-10.3, -5.2, 20.7, 18.1])        # clipped top-left
    boxes[1] = torch.tensor([100.5, 80.0, 140.2, 110.9])      # clipped bottom-right
    boxes[2] = torch.tensor([50.2, 40.7, 50.9, 41.1])         # sub-pixel box -> 1x1 .. 2x2 paste
    boxes[3] = torch.tensor([200.0, 200.0, 230.0, 230.0])     # fully outside
    out = ref.paste_masks_in_image(masks, boxes, (H, W), padding=1)
    _save(name, masks=_np(masks), boxes=_np(boxes), shape=np.array([H, W]), out=_np(out))


def make_roi(ref, name, seed):
    """Detect.multiscale_roi_align (yolo_head.py:279-299) on random features of a 3-level pyramid."""
    g = torch.Generator().manual_seed(seed)
    det = ref.Detect(ch=[8, 8, 8], anchors=synth.ANCHORS_3, strides=synth.STRIDES_3, nc=4, masks={}, is_scripting=True)
    tile, C, K = 160, 6, 28
    feats = [torch.randn((2, C, tile // s, tile // s), generator=g) for s in synth.STRIDES_3]
    c = torch.rand((K, 2), generator=g) * tile
    s = torch.rand((K, 2), generator=g) * 40 + 4
    b = torch.cat([c - s / 2, c + s / 2], 1)
    b[0] = torch.tensor([-30.0, -20.0, 12.5, 9.0])          # sticks out top-left
    b[1] = torch.tensor([150.0, 140.0, 200.0, 190.0])       # sticks out bottom-right
    b[2] = torch.tensor([300.0, 300.0, 340.0, 340.0])       # fully outside: every sample out of range
    b[3] = torch.tensor([40.2, 50.1, 40.4, 50.3])           # sub-pixel box (roi size clamps to 1)
    b[4] = torch.tensor([-5.0, -5.0, 165.0, 165.0])         # whole image and more
    boxes = torch.cat([torch.randint(0, 2, (K, 1), generator=g).float(), b], 1)
    levels = torch.randint(0, 3, (K,), generator=g).float()
    levels[5] = 3.0                                          # no such level: row stays zero
    out = det.multiscale_roi_align(feats, boxes, levels)
    arrays = {"boxes": _np(boxes), "levels": _np(levels), "out": _np(out),
              "strides": np.array(synth.STRIDES_3, dtype=np.float32)}
    for i, f in enumerate(feats):
        arrays[f"feat{i}"] = _np(f)
    _save(name, **arrays)


def make_match(ref, name, seed):
    """APMeter.add (metrics.py:270-303) over three images (one without predictions, one without ground truth)."""
    g = torch.Generator().manual_seed(seed)
    meter = ref.APMeter()
    arrays = {}
    for i, (k, n_gt) in enumerate([(60, 50), (0, 7), (25, 0), (80, 90)]):
        gc = torch.rand((n_gt, 2), generator=g) * 200
        gs = torch.rand((n_gt, 2), generator=g) * 24 + 10
        gt = torch.cat([gc - gs / 2, gc + gs / 2], 1)
        m = min(k, n_gt)
        # predictions: jittered copies of some ground-truth boxes + false positives
        pb = torch.cat([gt[:m] + torch.randn((m, 4), generator=g) * 2.0,
                        torch.rand((k - m, 4), generator=g) * 100 + torch.tensor([0., 0., 100., 100.])])
        if m > 2:
            pb[1] = gt[0]                                     # exact hit on another box's target (iou = 1)
        pb = pb[torch.randperm(k, generator=g)] if k else pb
        out = {'boxes': pb, 'scores': torch.rand(k, generator=g), 'labels': torch.randint(-1, 5, (k,), generator=g)}
        tgt = {'boxes': gt, 'labels': torch.randint(1, 5, (n_gt,), generator=g)}
        meter.add(out, tgt)
        for key, v in out.items():
            arrays[f"out{i}_{key}"] = _np(v)
        for key, v in tgt.items():
            arrays[f"tgt{i}_{key}"] = _np(v)
    arrays["n_images"] = np.int64(4)
    for f in ("scores", "y_pred", "y_true", "ious", "m_pred", "m_true"):
        arrays["meter_" + f] = _np(getattr(meter, f))
    arrays["meter_n"] = np.array([meter.n_pred, meter.n_true, meter.n_match], dtype=np.int64)
    _save(name, **arrays)


def make_scale_coords(ref, name, seed):
    """C1: scale_coords + clip_coords of the reference (utils_general.py:161-190), with and without ratio_pad."""
    g = torch.Generator().manual_seed(seed)
    arrays = {}
    cases = [((640, 640), (480, 720), None), ((640, 512), (1080, 810), None), ((320, 320), (100, 333), None),
             ((640, 640), (500, 375), ((1.28, 1.28), (80.0, 0.0)))]
    for i, (img1, img0, rp) in enumerate(cases):
        c = torch.rand((200, 4), generator=g) * torch.tensor([img1[1], img1[0], img1[1], img1[0]]) * 1.2 - 30.0
        out = ref.scale_coords(img1, c.clone(), img0, rp)
        arrays[f"in{i}"], arrays[f"out{i}"] = _np(c), _np(out)
        arrays[f"img1_{i}"], arrays[f"img0_{i}"] = np.array(img1, dtype=np.int64), np.array(img0, dtype=np.int64)
        arrays[f"rp{i}"] = np.array([rp[0][0], rp[1][0], rp[1][1]] if rp else [0, 0, 0], dtype=np.float64)
    arrays["n"] = np.int64(len(cases))
    _save(name, **arrays)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    make_detect(ref, "detect_640_l3", 640, synth.STRIDES_3, synth.ANCHORS_3, 4, seed=1, conf=0.001, max_det=300)
    make_detect(ref, "detect_320_l4", 320, synth.STRIDES_4, synth.ANCHORS_4, 7, seed=2, conf=0.002, max_det=1000)
    make_nms(ref, "nms_rows", seed=3)
    make_hier(ref, "hier_tree", seed=4)
    make_merge(ref, "tile_merge", seed=5)
    make_paste(ref, "paste_masks", seed=6)
    make_roi(ref, "roi_align", seed=7)
    make_match(ref, "ap_match", seed=8)
    make_scale_coords(ref, "scale_coords", seed=9)


if __name__ == "__main__":
    sys.exit(main())
