"""Synthetic head-output generator (seeded).  Shared by bench.py, the tests and the golden generator
bench.py so that oracle and kernels always see identical bytes."""
import math
from typing import List, Sequence, Tuple

import torch

# anchors of the reference's nuclei configs (metayolo/hub/yolov5l6-mask.yaml:7-11 uses 4 levels;
# the 3-level variant below is the P3-P5 subset named by BASELINE.json configs[1..2])
ANCHORS_3 = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
ANCHORS_4 = [[19, 27, 44, 40, 38, 94], [96, 68, 86, 152, 180, 137], [140, 301, 303, 264, 238, 542],
             [436, 615, 739, 380, 925, 792]]
STRIDES_3 = [8, 16, 32]
STRIDES_4 = [8, 16, 32, 64]


def level_shapes(tile: int, strides: Sequence[int]) -> List[Tuple[int, int]]:
    return [(tile // s, tile // s) for s in strides]


def nuclei_logits(bs: int, tile: int, nc: int, n_cand: int, seed: int, anchors=ANCHORS_3, strides=STRIDES_3,
                  extra: int = 0, conf: float = 0.25, size_range=(12.0, 36.0), generator_device="cpu"):
    """Raw head logits [bs,na,ny,nx,5+nc+extra] per level such that ~n_cand rows per tile pass
    sigmoid(obj) > conf and decode to nuclei-sized boxes (cfg 2/3).
    Objectness logits are N(mu,1) with mu chosen so the expected pass count is n_cand; the wh logits
    of every row are set so the decoded box has a side drawn from size_range."""
    g = torch.Generator(device=generator_device).manual_seed(seed)
    na = len(anchors[0]) // 2
    shapes = level_shapes(tile, strides)
    N = sum(na * ny * nx for ny, nx in shapes)
    frac = min(max(n_cand / N, 1e-6), 0.999)
    # P(N(mu,1) > logit(conf)) = frac
    z = math.sqrt(2.0) * torch.erfinv(torch.tensor(1.0 - 2.0 * frac, dtype=torch.float64)).item()  # Phi^-1(1-frac)
    mu = math.log(conf / (1.0 - conf)) - z
    dets = []
    for l, (ny, nx) in enumerate(shapes):
        no = 5 + nc + extra
        d = torch.randn((bs, na, ny, nx, no), generator=g, device=generator_device)
        d[..., 4] += mu
        a = torch.tensor(anchors[l], dtype=torch.float32, device=generator_device).view(na, 2)
        side = torch.rand((bs, na, ny, nx, 2), generator=g, device=generator_device) * (size_range[1] - size_range[0]) + size_range[0]
        # (2*sig)^2 * anchor = side  ->  sig = sqrt(side/anchor)/2, clamped into (0,1)
        sig = (side / a.view(1, na, 1, 1, 2)).sqrt().mul(0.5).clamp(0.02, 0.98)
        d[..., 2:4] = torch.log(sig / (1 - sig))
        dets.append(d.contiguous())
    return dets
