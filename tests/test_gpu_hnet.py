"""GPU parity tests of the hnet multi-level heads (H1-H3): C-ABI kernels vs torchvision's own functions, which is
what the reference's hnet/detection/mask_rcnn.py calls (:67 BoxCoder.decode, :72 filter_proposals, :192
postprocess_detections, :248 maskrcnn_inference).  Bars: boxes/scores within 1e-5 relative (expf / softmax differ by
ulps between libraries), kept index structure (which box, which class, order) identical."""
import math

import pytest
import torch

import hd_yolo_b200 as hdy
from hd_yolo_b200 import hnet, synth_hnet as sh
from oracle import port

pytestmark = pytest.mark.gpu
REL, ATOL = 1e-5, 1e-3   # px: coordinates are O(100-1000) so 1e-5 relative ~ 1e-3..1e-2 px; clip to 0 needs atol


def _close(a, b):
    return bool(((a.double() - b.double()).abs() <= REL * b.double().abs() + ATOL).all())


@pytest.mark.parametrize("weights", [(1., 1., 1., 1.), (10., 10., 5., 5.)])
def test_box_decode_vs_torchvision(cuda_device, weights):
    g = torch.Generator().manual_seed(0)
    R, C = 1000, 5
    c = torch.rand((R, 2), generator=g) * 1000
    wh = 4 + torch.rand((R, 2), generator=g) * 300
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    deltas = torch.randn((R, C * 4), generator=g) * 2.0
    deltas[0, 2] = 100.0          # exercises the bbox_xform_clip clamp
    ref = port.rcnn_box_decode(deltas, [boxes[:400], boxes[400:]], weights)
    got = hnet.box_decode(deltas.to(cuda_device), [boxes[:400].to(cuda_device), boxes[400:].to(cuda_device)], weights)
    assert got.shape == ref.shape == (R, C, 4)
    assert _close(got.cpu(), ref)


def _match_sets(gb, gs, rb, rs):
    """same survivors in the same order (scores descending), scores within tolerance.  Runs of EXACTLY equal scores
    may be permuted: torchvision's class-separated batched_nms orders its result with an unstable sort, ours breaks
    ties by the lower row index."""
    assert gb.shape == rb.shape, (gb.shape, rb.shape)
    gb, gs = gb.cpu(), gs.cpu()
    assert bool((gs[:-1] >= gs[1:]).all())

    def canon(b, s):
        key = torch.stack([-s.double(), b[:, 0].double(), b[:, 1].double(), b[:, 2].double(), b[:, 3].double()], 1)
        order = sorted(range(len(s)), key=lambda i: tuple(key[i].tolist()))
        return b[order], s[order]
    if not (_close(gb, rb) and _close(gs, rs)):
        gb, gs = canon(gb, gs)
        rb, rs = canon(rb, rs)
    assert _close(gb, rb) and _close(gs, rs)


@pytest.mark.parametrize("mode,n_img,size,pre,post", [("torchvision-cpu", 2, 256, 1000, 1000), ("vanilla", 3, 512, 600, 300),
                                                      ("torchvision-cpu", 1, 128, 200, 50)])
def test_rpn_filter_proposals_vs_torchvision(cuda_device, mode, n_img, size, pre, post):
    anchors, counts, obj, deltas = sh.rpn_inputs(n_img, size, seed=size + n_img)
    shapes = [(size, size - 16)] * n_img
    prop_ref = port.rcnn_box_decode(deltas, [anchors] * n_img).view(n_img, -1, 4)
    rb, rs = port.rpn_filter_proposals(prop_ref, obj, shapes, counts, pre, post, 0.7, 0.0)
    dev = cuda_device
    prop = hnet.box_decode(deltas.to(dev), anchors.to(dev), boxes_rows=anchors.shape[0]).view(n_img, -1, 4)
    assert _close(prop.cpu(), prop_ref)
    # decisions on identical inputs: feed the oracle's proposals (the decode is checked above)
    gb, gs = hnet.rpn_filter_proposals(prop_ref.to(dev), obj.to(dev), shapes, counts, pre, post, 0.7, 0.0, mode=mode)
    if mode == "vanilla":
        # torchvision takes the class-separated form itself above 4000 coordinates on CPU: same rule here
        assert all(min(c, pre) for c in counts) and sum(min(c, pre) for c in counts) * 4 > 4000
    for i in range(n_img):
        _match_sets(gb[i], gs[i], rb[i], rs[i])


@pytest.mark.parametrize("mode,n_img,rois,C,thr", [("torchvision-cpu", 2, 300, 5, 0.05), ("vanilla", 2, 1000, 8, 0.02),
                                                   ("torchvision-cpu", 3, 50, 3, 0.3)])
def test_roi_postprocess_detections_vs_torchvision(cuda_device, mode, n_img, rois, C, thr):
    size = 512
    props, cl, br = sh.roi_inputs(n_img, rois, C, size, seed=rois + C)
    shapes = [(size, size)] * n_img
    rb, rs, rl = port.roi_postprocess_detections(cl, br, props, shapes, score_thresh=thr)
    dev = cuda_device
    gb, gs, gl = hnet.roi_postprocess_detections(cl.to(dev), br.to(dev), [p.to(dev) for p in props], shapes,
                                                 score_thresh=thr, mode=mode)
    if mode == "vanilla":   # make sure torchvision itself was in its class-separated regime for every image
        from torchvision.ops import boxes as box_ops
        assert rois * (C - 1) * 4 > 4000
    for i in range(n_img):
        _match_sets(gb[i], gs[i], rb[i], rs[i])
        assert torch.equal(gl[i].cpu(), rl[i])


def test_maskrcnn_inference_vs_torchvision(cuda_device):
    g = torch.Generator().manual_seed(4)
    K, C, M = 37, 5, 28
    x = torch.randn((K, C, M, M), generator=g)
    labels = [torch.randint(1, C, (20,), generator=g), torch.randint(1, C, (17,), generator=g)]
    ref = port.maskrcnn_inference(x, labels)
    got = hnet.maskrcnn_inference(x.to(cuda_device), [l.to(cuda_device) for l in labels])
    for a, b in zip(got, ref):
        assert a.shape == b.shape and _close(a.cpu(), b)


def test_cross_level_merge_vs_oracle(cuda_device):
    """10x structure detections (scale 4 -> 40x frame) + 40x nuclei tiles merged by Ensemble.merge."""
    g = torch.Generator().manual_seed(7)

    def tiles(n_tiles, k, span, lo, hi, origin_step):
        out = []
        for t in range(n_tiles):
            c = torch.rand((k, 2), generator=g) * span
            wh = lo + torch.rand((k, 2), generator=g) * (hi - lo)
            out.append({'boxes': torch.cat([c - wh / 2, c + wh / 2], 1), 'scores': torch.rand((k,), generator=g),
                        'labels': torch.randint(1, 4, (k,), generator=g),
                        'roi': torch.tensor([t * origin_step, 0., t * origin_step + span, span])})
        return out
    lv40 = tiles(4, 400, 1024., 12., 36., 960.)
    lv10 = tiles(2, 60, 1024., 20., 120., 960.)
    params = {'conf_thres': 0.2, 'iou_thres': 0.45, 'max_det': 100000}
    ref_parts = []
    for scale, tl in ((1.0, lv40), (4.0, lv10)):
        m = port.merge_outputs([{k: (v.clone() if torch.is_tensor(v) else v) for k, v in t.items()} for t in tl])
        port.rescale_outputs(m, scale)
        ref_parts.append({'det': m})
    ref = port.ensemble_merge(ref_parts, params)['det']
    dev = cuda_device
    to = lambda tl: [{k: v.to(dev) for k, v in t.items()} for t in tl]
    got = hnet.cross_level_merge([(1.0, to(lv40)), (4.0, to(lv10))], params)
    assert torch.equal(got['boxes'].cpu(), ref['boxes'])
    assert torch.equal(got['scores'].cpu(), ref['scores'])
    assert torch.equal(got['labels'].cpu(), ref['labels'])
