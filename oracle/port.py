"""ORACLE -- test infrastructure, not product code.

CPU restatement of impromptuRong/hd_yolo's inference post-processing path, written from the
reference's behaviour (file:line cited per function, relative to the reference root).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module; nothing under ``hd_yolo_b200/`` does.

The reference is Python on torch + torchvision, so the restatement uses the same CPU library ops
for the third-party arithmetic (``torch.sigmoid``, ``torchvision.ops.nms``,
``F.interpolate``) -- torch 2.11 / torchvision 0.26, unpinned by the reference (it ships no
requirements file).  ``oracle/nms_core.c`` restates torchvision's greedy NMS in plain C so that
third-party op is pinned independently as well (tests/test_oracle.py cross-checks the two).

PINNING: the reference has no tests, golden vectors or fixtures for this path.  This port is
pinned against outputs of the reference itself, generated in the build container by
``oracle/make_golden.py`` (which imports /root/reference through the shim in
``oracle/ref_shim.py``) and committed under ``tests/golden/``.  ``process_mask`` (upstream
ultralytics/yolov5 v7.0 ``utils/segment/general.py``, not vendored by the reference) and the hnet
cross-level merge (a TODO in the reference, hnet/hnet_new.py:275) have no reference output to pin
against: "parity unpinned" for those two, see DESIGN.md.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
import torchvision


# --------------------------------------------------------------------------------------- decode
def anchor_grids(anchors: Sequence[Sequence[float]], strides: Sequence[float]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Detect.__init__ (yolo_head.py:58-59): anchors are stored as anchors/stride; _make_grid
    (:427) multiplies them back.  Returns (strides [nl], anchor_grid [nl,na,2])."""
    s = torch.tensor(list(strides)).float()
    a = torch.tensor(anchors).float().view(len(anchors), -1, 2) / s.view(-1, 1, 1)
    return s, a * s.view(-1, 1, 1)


def make_grid(nx: int, ny: int) -> torch.Tensor:
    """Detect._make_grid (yolo_head.py:419-429): grid[y, x] = (x, y)."""
    y, x = torch.arange(ny, dtype=torch.float32), torch.arange(nx, dtype=torch.float32)
    yv, xv = torch.meshgrid(y, x, indexing='ij')
    return torch.stack((xv, yv), 2)


def compute_proposals(dets: List[torch.Tensor], anchors, strides) -> List[torch.Tensor]:
    """Detect.compute_proposals (yolo_head.py:185-213), non-scripting branch (:206-210); the
    in-place scripting branch (:203-204) performs the same arithmetic."""
    s, ag = anchor_grids(anchors, strides)
    preds = []
    for i, det in enumerate(dets):
        y = det.sigmoid()
        bs, na, ny, nx, no = y.shape
        grid = make_grid(nx, ny).to(det.device).expand(1, na, ny, nx, 2)      # buffers live on the model's device
        anchor_grid = ag[i].to(det.device).view(1, na, 1, 1, 2).expand(1, na, ny, nx, 2)
        xy, wh, conf = y.tensor_split((2, 4), -1)
        xy = (xy * 2. - 0.5 + grid) * s[i]
        wh = (wh * 2.) ** 2 * anchor_grid
        preds.append(torch.cat((xy, wh, conf), -1))
    return preds


def concat_levels(preds: List[torch.Tensor]) -> torch.Tensor:
    """Level-id pad + concat at the top of Detect.compute_outputs (yolo_head.py:311-312)."""
    no = preds[0].shape[-1]
    return torch.cat([F.pad(y.view(y.shape[0], -1, no), [0, 1], value=float(idx))
                      for idx, y in enumerate(preds)], 1)


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """utils_general.py:121-128."""
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


# ------------------------------------------------------------------------------------------ NMS
def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """The third-party op the reference calls (utils_general.py:342, :507; yolo.py:195)."""
    return torchvision.ops.nms(boxes, scores, iou_thres)


def nms_per_image(preds: torch.Tensor, nc: int, conf_thres: float = 0.25, iou_thres: float = 0.45,
                  max_det: int = 300) -> List[Dict[str, torch.Tensor]]:
    """utils_general.py:299-356 without the 10 s wall-clock exit (:351-354)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    outputs = []
    for x in preds:
        boxes = xywh2xyxy(x[:, :4])
        scores = x[:, 4:4 + 1 + nc]
        extra = x[:, 5 + nc:]
        keep = torchvision.ops.remove_small_boxes(boxes, min_size=2.)          # :332
        boxes, scores, extra = boxes[keep], scores[keep], extra[keep]
        nms_scores = scores[:, 0]
        keep = nms_scores > conf_thres                                         # :336-337
        boxes, scores, extra, nms_scores = boxes[keep], scores[keep], extra[keep], nms_scores[keep]
        if len(boxes):
            keep = nms(boxes, nms_scores, iou_thres)[:max_det]                 # :342
            boxes, scores, extra = boxes[keep], scores[keep], extra[keep]
        outputs.append({'boxes': boxes, 'scores': scores, 'extra': extra})
    return outputs


def non_max_suppression(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None,
                        agnostic=False, multi_label=False, labels=(), max_det=300) -> List[torch.Tensor]:
    """utils_general.py:423-523 (merge=False branch; no wall-clock exit).  Does not mutate its input
    (the reference only mutates its own boolean-indexed copy, :460,476)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    bs = prediction.shape[0]
    nc = prediction.shape[2] - 5
    xc = prediction[..., 4] > conf_thres
    max_wh, max_nms = 7680, 30000
    multi_label &= nc > 1
    output = [torch.zeros((0, 6))] * bs
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if labels and len(labels[xi]):
            lb = labels[xi]
            v = torch.zeros((len(lb), nc + 5))
            v[:, :4] = lb[:, 1:5]
            v[:, 4] = 1.0
            v[range(len(lb)), lb[:, 0].long() + 5] = 1.0
            x = torch.cat((x, v), 0)
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]
        box = xywh2xyxy(x[:, :4])
        if multi_label:
            i, j = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[i], x[i, j + 5, None], j[:, None].float()), 1)
        else:
            conf, j = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        boxes, scores = x[:, :4] + c, x[:, 4]
        i = nms(boxes, scores, iou_thres)
        if i.shape[0] > max_det:
            i = i[:max_det]
        output[xi] = x[i]
    return output


# --------------------------------------------------------------------------- score / label select
def default_descendants(nc: int) -> Dict[int, List[int]]:
    """Detect.build_hierarchical_tree default {0: {1..nc}} -> get_descendants (yolo_head.py:481-511)."""
    return {0: list(range(1, nc + 1))}


def hierarchical_scores(x: torch.Tensor, descendants: Dict[int, List[int]]) -> torch.Tensor:
    """yolo_head.py:473-479 (in place)."""
    for k, v in descendants.items():
        x[:, v] *= x[:, k:k + 1]
    return x


def select_scores(scores: torch.Tensor, conf_thres: float, descendants: Dict[int, List[int]],
                  multi_label: bool = False):
    """Detect.compute_outputs score/label block (yolo_head.py:336-345).  Returns (scores, labels)."""
    scores = hierarchical_scores(scores, descendants)
    if multi_label:
        return scores, scores > conf_thres
    obj_scores = scores[..., 0]
    cls_scores, cls_labels = scores[..., 1:].max(1)
    out_scores = torch.where(cls_scores > conf_thres, cls_scores, obj_scores)
    labels = torch.where(cls_scores > conf_thres, cls_labels + 1, -100)
    return out_scores, labels


def compute_outputs(preds: List[torch.Tensor], nc: int, nms_params: Dict[str, float],
                    descendants: Optional[Dict[int, List[int]]] = None, multi_label: bool = False):
    """Detect.compute_outputs with compute_masks=False (yolo_head.py:301-355)."""
    cat = concat_levels(preds)
    outputs = nms_per_image(cat, nc=nc, conf_thres=nms_params['conf_thres'], iou_thres=nms_params['iou_thres'],
                            max_det=int(nms_params['max_det']))
    desc = default_descendants(nc) if descendants is None else descendants
    results = []
    for o in outputs:
        s, l = select_scores(o['scores'], nms_params['conf_thres'], desc, multi_label)
        results.append({'boxes': o['boxes'], 'scores': s, 'labels': l, 'extra': o['extra']})
    return results


# ------------------------------------------------------------------------------------------ masks
def mask_select(mask_logits: torch.Tensor, labels: torch.Tensor, mask_indices: torch.Tensor) -> torch.Tensor:
    """Mask tail of compute_outputs (yolo_head.py:332, 346-353) for one image, with the float clamp
    at :348 (which raises IndexError on torch >= 2) restated as the integer clamp it intends."""
    m = mask_logits.sigmoid()
    mask_labels = mask_indices[labels.clamp(min=0)]
    index = torch.arange(len(m))
    masks = m[index, mask_labels.clamp(min=0)][:, None]
    masks[mask_labels < 0] = 0
    return masks


def paste_masks_in_image(masks: torch.Tensor, boxes: torch.Tensor, img_shape: Tuple[int, int],
                         padding: int = 1) -> torch.Tensor:
    """torchvision.models.detection.roi_heads.paste_masks_in_image as called by val_nuclei.py:169-176 and
    evaluation.py:122-123: pad the MxM mask by `padding`, grow the box by (M+2p)/M about its centre,
    truncate to int64, bilinear-resize (align_corners=False) to (h, w) = box extent + 1 (>= 1), paste
    the in-image part into a zero [H, W] canvas."""
    M = masks.shape[-1]
    scale = float(M + 2 * padding) / M
    padded = F.pad(masks, (padding,) * 4)
    w_half = (boxes[:, 2] - boxes[:, 0]) * 0.5
    h_half = (boxes[:, 3] - boxes[:, 1]) * 0.5
    x_c = (boxes[:, 2] + boxes[:, 0]) * 0.5
    y_c = (boxes[:, 3] + boxes[:, 1]) * 0.5
    w_half = w_half * scale
    h_half = h_half * scale
    bexp = torch.stack([x_c - w_half, y_c - h_half, x_c + w_half, y_c + h_half], 1).to(torch.int64)
    im_h, im_w = img_shape
    res = []
    for m, b in zip(padded, bexp):
        x0b, y0b, x1b, y1b = [int(v) for v in b]
        w = max(x1b - x0b + 1, 1)
        h = max(y1b - y0b + 1, 1)
        r = F.interpolate(m[None], size=(h, w), mode='bilinear', align_corners=False)[0, 0]
        canvas = torch.zeros((im_h, im_w), dtype=r.dtype)
        x_0, x_1 = max(x0b, 0), min(x1b + 1, im_w)
        y_0, y_1 = max(y0b, 0), min(y1b + 1, im_h)
        if x_1 > x_0 and y_1 > y_0:
            canvas[y_0:y_1, x_0:x_1] = r[(y_0 - y0b):(y_1 - y0b), (x_0 - x0b):(x_1 - x0b)]
        res.append(canvas)
    if res:
        return torch.stack(res, 0)[:, None]
    return masks.new_empty((0, 1, im_h, im_w))


def crop_mask(masks: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """ultralytics/yolov5 v7.0 utils/segment/general.py::crop_mask (third-party, not vendored):
    zero everything outside x1 <= col < x2, y1 <= row < y2."""
    n, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, device=masks.device, dtype=x1.dtype)[None, None, :]   # upstream: device=masks.device
    c = torch.arange(h, device=masks.device, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def process_mask(protos: torch.Tensor, masks_in: torch.Tensor, bboxes: torch.Tensor, shape: Tuple[int, int],
                 upsample: bool = False) -> torch.Tensor:
    """ultralytics/yolov5 v7.0 utils/segment/general.py::process_mask (third-party, not vendored; the
    reference contains no such function -- parity unpinned).
    protos [c, mh, mw], masks_in [n, c] (after NMS), bboxes [n, 4] in image pixels, shape (ih, iw)."""
    c, mh, mw = protos.shape
    ih, iw = shape
    masks = (masks_in @ protos.float().view(c, -1)).sigmoid().view(-1, mh, mw)
    db = bboxes.clone()
    db[:, 0] *= mw / iw
    db[:, 2] *= mw / iw
    db[:, 3] *= mh / ih
    db[:, 1] *= mh / ih
    masks = crop_mask(masks, db)
    if upsample:
        masks = F.interpolate(masks[None], shape, mode='bilinear', align_corners=False)[0]
    return masks.gt_(0.5)


# ------------------------------------------------------------------------------------ coordinates
def clip_coords(boxes: torch.Tensor, shape) -> None:
    """utils_general.py:181-190 (tensor branch, in place)."""
    boxes[:, 0].clamp_(0, shape[1])
    boxes[:, 1].clamp_(0, shape[0])
    boxes[:, 2].clamp_(0, shape[1])
    boxes[:, 3].clamp_(0, shape[0])


def scale_coords(img1_shape, coords: torch.Tensor, img0_shape, ratio_pad=None) -> torch.Tensor:
    """utils_general.py:161-178 (in place on coords)."""
    if isinstance(img1_shape, int):
        img1_shape = (img1_shape, img1_shape)
    if isinstance(img0_shape, int):
        img0_shape = (img0_shape, img0_shape)
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    clip_coords(coords, img0_shape)
    return coords


# ------------------------------------------------------------------------------- tiling and merge
def sliding_window_scanner(image_size, roi_size=None, overlap=0) -> torch.Tensor:
    """hnet/utils.py:37-62 (== metayolo/models/utils_o.py:37-62): tile origins every roi-overlap
    pixels, x fastest, boxes clipped to the image (edge tiles are slivers)."""
    if roi_size is None:
        return torch.tensor([[0., 0., image_size[0], image_size[1]]], dtype=torch.float32)
    pair = lambda v: (v, v) if isinstance(v, (int, float)) else v
    h, w = pair(image_size)
    roi_h, roi_w = pair(roi_size)
    x0 = torch.arange(0, w, roi_w - overlap, dtype=torch.float32) if w > roi_w else torch.zeros((1,))
    y0 = torch.arange(0, h, roi_h - overlap, dtype=torch.float32) if h > roi_h else torch.zeros((1,))
    y0, x0 = torch.meshgrid(y0, x0, indexing='ij')
    x0, y0 = x0.reshape(-1), y0.reshape(-1)
    boxes = torch.stack((x0, y0, x0 + roi_w, y0 + roi_h), dim=1)
    return torchvision.ops.boxes.clip_boxes_to_image(boxes, (h, w))


def split_by_sizes(x, sizes):
    """hnet/utils.py:21-26."""
    assert len(x) == sum(sizes)
    ends = list(itertools.accumulate(list(sizes)))
    starts = [0] + list(ends[:-1])
    return [x[s:e] for s, e in zip(starts, ends)]


def merge_outputs(r: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Detect.merge_outputs (yolo_head.py:450-463): shift boxes by the tile origin, concatenate; no NMS."""
    boxes = torch.cat([_['boxes'] + _['boxes'].new([_['roi'][0], _['roi'][1], _['roi'][0], _['roi'][1]]) for _ in r])
    res = {'boxes': boxes, 'labels': torch.cat([_['labels'] for _ in r]), 'scores': torch.cat([_['scores'] for _ in r])}
    if 'masks' in r[0]:
        res['masks'] = torch.cat([_['masks'] for _ in r])
    return res


def rescale_outputs(r: Dict[str, torch.Tensor], scale: float = 1.0):
    """Detect.rescale_outputs (yolo_head.py:465-471), in place."""
    if scale != 1.0:
        r['boxes'] *= scale
    return r


def ensemble_merge(x: List[Dict[str, Dict[str, torch.Tensor]]], nms_params: Dict[str, float]):
    """Ensemble.merge (yolo.py:165-204): per task, concatenate, keep scores > conf, class-agnostic
    NMS on the final scores, first max_det."""
    task_ids = set().union(*x)
    res = {}
    for task_id in task_ids:
        parts = [r[task_id] for r in x if task_id in r]
        boxes = torch.cat([p['boxes'] for p in parts])
        scores = torch.cat([p['scores'] for p in parts])
        labels = torch.cat([p['labels'] for p in parts])
        masks = None
        if any('masks' in p for p in parts):
            ref = [p['masks'] for p in parts if 'masks' in p][0]
            masks = torch.cat([p['masks'] if 'masks' in p else torch.zeros(ref.shape[1:]).to(ref.device, ref.dtype)
                               for p in parts])
        keep = scores > nms_params['conf_thres']
        boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
        if masks is not None:
            masks = masks[keep]
        if len(boxes):
            keep = nms(boxes, scores, nms_params['iou_thres'])[:int(nms_params['max_det'])]
            boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
            if masks is not None:
                masks = masks[keep]
        res[task_id] = {'boxes': boxes, 'scores': scores, 'labels': labels}
        if masks is not None:
            res[task_id]['masks'] = masks
    return res


def project_roi_results_on_image(results, rois, image_shape=None):
    """hnet/detection/utils_det.py:143-160: shift boxes by the roi origin, optional clip, concatenate."""
    out: Dict[str, List[torch.Tensor]] = {}
    for r, roi in zip(results, rois):
        x0, y0 = roi[0], roi[1]
        b = r['boxes'].clone()
        b[:, 0] += x0
        b[:, 1] += y0
        b[:, 2] += x0
        b[:, 3] += y0
        if image_shape is not None:
            b = torchvision.ops.boxes.clip_boxes_to_image(b, (image_shape[0], image_shape[1]))
        for k, v in r.items():
            out.setdefault(k, []).append(b if k == 'boxes' else v)
    return {k: torch.cat(v) for k, v in out.items()}


# --------------------------------------------------------------------------------------- hnet heads (H1-H3)
# hnet/detection/mask_rcnn.py hands all of this arithmetic to torchvision (mask_rcnn.py:67, :72, :192, :248; thresholds
# hnet/detection/utils_det.py:16-52).  `import hnet.detection` fails in the reference itself (NameError: tmdet,
# utils_det.py:220), so the oracle calls the same torchvision (0.26, unpinned) functions directly.
def rcnn_box_decode(rel_codes: torch.Tensor, boxes: List[torch.Tensor], weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    from torchvision.models.detection._utils import BoxCoder
    return BoxCoder(tuple(float(w) for w in weights)).decode(rel_codes, boxes)


def rpn_filter_proposals(proposals, objectness, image_shapes, num_anchors_per_level, pre_nms_top_n=1000,
                         post_nms_top_n=1000, nms_thresh=0.7, score_thresh=0.0):
    from torchvision.models.detection.rpn import RegionProposalNetwork
    rpn = RegionProposalNetwork(None, None, 0.7, 0.3, 256, 0.5, dict(training=pre_nms_top_n, testing=pre_nms_top_n),
                                dict(training=post_nms_top_n, testing=post_nms_top_n), nms_thresh, score_thresh)
    rpn.eval()
    return rpn.filter_proposals(proposals, objectness, image_shapes, num_anchors_per_level)


def roi_postprocess_detections(class_logits, box_regression, proposals, image_shapes, box_weights=(10., 10., 5., 5.),
                               score_thresh=0.05, nms_thresh=0.5, detections_per_img=100):
    from torchvision.models.detection.roi_heads import RoIHeads
    rh = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, box_weights, score_thresh, nms_thresh, detections_per_img)
    rh.eval()
    return rh.postprocess_detections(class_logits, box_regression, proposals, image_shapes)


def maskrcnn_inference(x: torch.Tensor, labels: List[torch.Tensor]) -> List[torch.Tensor]:
    from torchvision.models.detection.roi_heads import maskrcnn_inference as f
    return f(x, labels)


# ------------------------------------------------------------------- next rows: RoIAlign, AP matching
def multiscale_roi_align(features: List[torch.Tensor], boxes: torch.Tensor, levels: torch.Tensor,
                         strides: Sequence[float], output_size: int = 14, sampling_ratio: int = 2,
                         aligned: bool = False) -> torch.Tensor:
    """Detect.multiscale_roi_align (metayolo/models/yolo_head.py:279-299): per level, torchvision.ops.roi_align of the
    boxes routed to it (spatial_scale = 1/stride, sampling_ratio = 2, aligned = ROI_ALIGN = False, :15), scattered
    into a zero [K, C, M, M] tensor.  boxes [K, 5] = (image index, xyxy)."""
    K, C, M = len(boxes), features[0].shape[1], output_size
    result = torch.zeros((K, C, M, M), dtype=features[0].dtype)
    for i, stride in enumerate(strides):
        idx = torch.where(levels == i)[0]
        fmap = torchvision.ops.roi_align(features[i], boxes[idx], (M, M), spatial_scale=1 / float(stride),
                                         sampling_ratio=sampling_ratio, aligned=aligned)
        result[idx] = fmap.to(result.dtype)
    return result


def box_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """box_iou (metayolo/models/utils_general.py:247-265): inter = clamp(min(a2,b2) - max(a1,b1), 0).prod(2);
    iou = inter / (area1[:, None] + area2 - inter), area = (x2-x1)*(y2-y1)."""
    a1, a2 = box1[:, None, :2], box1[:, None, 2:]
    b1, b2 = box2[:, :2], box2[:, 2:]
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp(0).prod(2)
    area1 = (box1[:, 2] - box1[:, 0]) * (box1[:, 3] - box1[:, 1])
    area2 = (box2[:, 2] - box2[:, 0]) * (box2[:, 3] - box2[:, 1])
    return inter / (area1[:, None] + area2 - inter)


def apmeter_match(output: Dict[str, torch.Tensor], target: Dict[str, torch.Tensor], iou_min: float = 0.5):
    """The matching half of APMeter.add (metayolo/models/metrics.py:271-284): predictions by score (descending), dense
    box_iou against the ground truth, pairs with iou >= iouv.min() sorted by IoU (descending).  Returns o_scores,
    o_labels, pred_idx (rank in score order), true_idx, ious.  Ties: both sorts are stable here (the reference calls
    torch.sort without `stable`, whose CPU kernel is stable)."""
    o_scores, order = torch.sort(output['scores'], descending=True, stable=True)
    o_labels = output['labels'][order]
    ious = box_iou(output['boxes'][order], target['boxes'])
    pred_idx, true_idx = torch.where(ious >= iou_min)
    ious, o2 = torch.sort(ious[pred_idx, true_idx], descending=True, stable=True)
    return o_scores, o_labels, pred_idx[o2], true_idx[o2], ious


class APMeterState:
    """Accumulator fields of APMeter (metrics.py:250-303) filled through apmeter_match."""

    def __init__(self):
        self.n_pred = self.n_true = self.n_match = 0
        self.scores, self.ious = torch.empty(0), torch.empty(0)
        self.y_pred = torch.empty(0, dtype=torch.int64)
        self.y_true = torch.empty(0, dtype=torch.int64)
        self.m_pred = torch.empty(0, dtype=torch.int64)
        self.m_true = torch.empty(0, dtype=torch.int64)

    def add(self, output, target, iou_min: float = 0.5):
        s, l, p, t, i = apmeter_match(output, target, iou_min)
        self.m_pred = torch.cat([self.m_pred, p + self.n_pred])
        self.m_true = torch.cat([self.m_true, t + self.n_true])
        self.ious = torch.cat([self.ious, i])
        self.n_match += len(i)
        self.y_true = torch.cat([self.y_true, target['labels']])
        self.n_true += len(target['boxes'])
        self.y_pred = torch.cat([self.y_pred, l])
        self.scores = torch.cat([self.scores, s])
        self.n_pred += len(output['boxes'])
