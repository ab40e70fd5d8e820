"""Synthetic inputs for the hnet multi-level heads (cfg 5): anchors over a feature pyramid,
RPN objectness logits and box deltas, RoI class logits and per-class box regression.  Seeded, generated on CPU."""
from typing import List, Tuple

import torch


def pyramid_anchors(image_size: int, strides=(4, 8, 16, 32, 64), sizes=(32, 64, 128, 256, 512),
                    ratios=(0.5, 1.0, 2.0)) -> Tuple[torch.Tensor, List[int]]:
    """torchvision AnchorGenerator layout: per level, per cell (row-major), per ratio.  Returns ([A,4], per-level A)."""
    out, counts = [], []
    for s, sz in zip(strides, sizes):
        n = max(image_size // s, 1)
        r = torch.tensor(ratios)
        hr, wr = torch.sqrt(r), 1.0 / torch.sqrt(r)
        base = (torch.stack([-wr * sz, -hr * sz, wr * sz, hr * sz], 1) / 2).round()
        sh = torch.arange(n, dtype=torch.float32) * s
        yy, xx = torch.meshgrid(sh, sh, indexing='ij')
        shift = torch.stack([xx.flatten(), yy.flatten(), xx.flatten(), yy.flatten()], 1)
        a = (shift[:, None, :] + base[None]).reshape(-1, 4)
        out.append(a)
        counts.append(a.shape[0])
    return torch.cat(out).contiguous(), counts


def rpn_inputs(n_img: int, image_size: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    anchors, counts = pyramid_anchors(image_size)
    A = anchors.shape[0]
    objectness = torch.randn((n_img, A), generator=g) * 2.0
    deltas = torch.randn((n_img * A, 4), generator=g) * 0.1
    return anchors, counts, objectness, deltas


def roi_inputs(n_img: int, rois_per_img: int, n_classes: int, image_size: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    props = []
    for _ in range(n_img):
        c = torch.rand((rois_per_img, 2), generator=g) * image_size
        wh = 16 + torch.rand((rois_per_img, 2), generator=g) * 200
        props.append(torch.cat([c - wh / 2, c + wh / 2], 1).contiguous())
    R = n_img * rois_per_img
    class_logits = torch.randn((R, n_classes), generator=g) * 2.0
    box_regression = torch.randn((R, n_classes * 4), generator=g) * 0.5
    return props, class_logits, box_regression
