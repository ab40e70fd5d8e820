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


# ------------------------------------------------------------------------------------------------ whole slide
def _tile_seed(seed: int, tile_index: int) -> int:
    return (int(seed) * 1000003 + int(tile_index) * 7919 + 12345) & 0x7fffffffffff


def slide_tile_logits(rois: torch.Tensor, tile: int, nc: int, seed: int, pitch: float = 18.7, extra: int = 0,
                      conf: float = 0.25, size_range=(12.0, 30.0), strides=STRIDES_3, anchors=ANCHORS_3,
                      device="cuda", background_mu: float = -6.0, first_tile: int = 0, dtype=torch.float32):
    """Raw head logits for a batch of slide tiles (rois [bs,4] = x0,y0,x1,y1 from sliding_window_scanner) cut from
    ONE global nuclei field (cfg 4): nuclei sit on a jittered grid of `pitch` px over the whole slide
    (~3 000 per 1024^2 tile); a nucleus is a pure function of its grid index, so the tiles that share an overlap band
    report the same nucleus with the same box (up to fp32 rounding of the tile-local decode) and a tile-dependent
    score jitter -- consistent duplicates for the merge NMS.  Every nucleus is written into the stride-8 level at the
    cell holding its centre, anchor chosen by size; all other rows are background (objectness logit ~ N(mu,1), far
    below the threshold).  Edge tiles (clipped windows) simply contain fewer nuclei.

    DETERMINISTIC AND RANK-INDEPENDENT: tile b of the batch is global tile `first_tile + b`, and everything random about
    it (background noise, score jitter, the `extra` coefficient channels) comes from a generator seeded with
    (seed, global tile index) alone -- however the tiles are batched or sharded over ranks, a tile's logits are the
    same bytes.  Two nuclei landing in the same (anchor, cell) are resolved by a fixed rule (the first in grid order
    wins), never by the order in which a scatter happens to commit duplicates.
    Returns the level tensors [bs,na,ny,nx,5+nc+extra] (`dtype`: float32, or float16 = what a half() model emits)."""
    bs = rois.shape[0]
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    na = len(anchors[0]) // 2
    no = 5 + nc + extra
    shapes = level_shapes(tile, strides)
    dets = [torch.empty((bs, na, ny, nx, no), dtype=torch.float32, device=dev) for (ny, nx) in shapes]
    rois = rois.to(dev, torch.float32)
    rois_h = rois.cpu().tolist()
    s0 = float(strides[0])
    a0 = torch.tensor(anchors[0], dtype=torch.float32, device=dev).view(na, 2)
    ncell = int(math.ceil(tile / pitch)) + 1
    ny0, nx0 = shapes[0]
    for b in range(bs):
        g.manual_seed(_tile_seed(seed, first_tile + b))
        for d in dets:
            d[b] = torch.randn(tuple(d.shape[1:]), generator=g, device=dev)
            d[b, ..., 4] += background_mu
        x0, y0, x1, y1 = rois_h[b]
        i0, j0 = int(math.floor(x0 / pitch)), int(math.floor(y0 / pitch))
        ii = torch.arange(i0, i0 + ncell, device=dev, dtype=torch.int64)
        jj = torch.arange(j0, j0 + ncell, device=dev, dtype=torch.int64)
        J, I = torch.meshgrid(jj, ii, indexing='ij')
        # hash (I, J) -> four uniforms in [0,1): pure function of the grid index
        h = (I * 73856093) ^ (J * 19349663)

        def u(k):
            v = (h * (2 * k + 1) + 0x9E3779B9 * (k + 1)) & 0x7fffffff
            v = (v * 1103515245 + 12345) & 0x7fffffff
            return (v >> 7).to(torch.float32) / float(1 << 24)
        cx = (I.float() + 0.15 + 0.7 * u(0)) * pitch
        cy = (J.float() + 0.15 + 0.7 * u(1)) * pitch
        side = size_range[0] + (size_range[1] - size_range[0]) * u(2)
        score = 0.3 + 0.65 * u(3)
        inside = (cx >= x0) & (cx < x1) & (cy >= y0) & (cy < y1)
        cx, cy, side, score = cx[inside] - x0, cy[inside] - y0, side[inside], score[inside]
        jitter = torch.rand((ncell * ncell,), generator=g, device=dev)[:score.numel()]   # fixed draw count per tile
        if cx.numel() == 0:
            continue
        gx = torch.clamp((cx / s0).floor().long(), 0, nx0 - 1)
        gy = torch.clamp((cy / s0).floor().long(), 0, ny0 - 1)
        a = torch.where(side < 14.0, 0, torch.where(side < 23.0, 1, 2)).long()
        sx = ((cx / s0 - gx.float()) + 0.5) / 2.0
        sy = ((cy / s0 - gy.float()) + 0.5) / 2.0
        sw = (side / a0[a, 0]).sqrt() / 2.0
        sh = (side / a0[a, 1]).sqrt() / 2.0
        sc = (score + 0.04 * (jitter - 0.5)).clamp(conf + 0.02, 0.99)
        sig = torch.stack([sx, sy, sw, sh, sc], 1).clamp(0.02, 0.98)
        # one nucleus per (anchor, cell): the first in grid order wins (stable sort + first of every run)
        lin = (a * ny0 + gy) * nx0 + gx
        order = torch.argsort(lin, stable=True)
        ls = lin[order]
        first = torch.ones_like(ls, dtype=torch.bool)
        first[1:] = ls[1:] != ls[:-1]
        win = order[first]
        dets[0][b, a[win], gy[win], gx[win], :5] = torch.log(sig[win] / (1 - sig[win]))
    if dtype != torch.float32:
        dets = [d.to(dtype) for d in dets]
    return [d.contiguous() for d in dets]


def slide_tile_protos(n: int, tile: int, seed: int, first_tile: int = 0, nm: int = 32, device="cuda",
                      dtype=torch.float32) -> torch.Tensor:
    """Prototype maps [n, nm, tile/4, tile/4] of the global tiles [first_tile, first_tile + n): N(0,1), seeded per
    global tile index like slide_tile_logits (the same bytes however the tiles are batched or sharded)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    mh = tile // 4
    out = torch.empty((n, nm, mh, mh), dtype=dtype, device=dev)
    for b in range(n):
        g.manual_seed(_tile_seed(seed + 77, first_tile + b))
        out[b] = torch.randn((nm, mh, mh), generator=g, device=dev).to(dtype)
    return out
