"""TEST INFRASTRUCTURE: a synthetic nuclei field detected by overlapping horizontal bands (one band per rank), so that
detections in the overlap are consistent duplicates (same nucleus, box jitter <= 1 px, score jitter)."""
import torch


def banded_detections(world, seed=0, width=2000.0, band=600.0, overlap=64.0, n_nuclei=1500, dup_in_band=0.1):
    """Returns parts = [(boxes [n_r,4], scores [n_r])] per rank (fp32, CPU) in slide coordinates."""
    g = torch.Generator().manual_seed(seed)
    H = band * world - overlap * (world - 1)
    c = torch.rand((n_nuclei, 2), generator=g) * torch.tensor([width, H])
    sz = 12.0 + 24.0 * torch.rand((n_nuclei, 2), generator=g)
    base_score = 0.2 + 0.8 * torch.rand((n_nuclei,), generator=g)
    parts = []
    for r in range(world):
        y0 = r * (band - overlap)
        y1 = y0 + band
        inside = (c[:, 1] >= y0) & (c[:, 1] < y1)
        idx = torch.nonzero(inside).flatten()
        # a few nuclei are reported twice by the same band (what a per-tile NMS at a lower threshold would leave)
        extra = idx[torch.rand((len(idx),), generator=g) < dup_in_band]
        idx = torch.cat([idx, extra])
        jit = torch.rand((len(idx), 4), generator=g) * 2 - 1
        half = sz[idx] / 2
        b = torch.cat([c[idx] - half, c[idx] + half], 1) + jit
        s = (base_score[idx] + 0.02 * (torch.rand((len(idx),), generator=g) - 0.5)).clamp(0.01, 0.999)
        # some exact score ties across ranks (tie order = global index)
        s[::17] = 0.5
        parts.append((b.float().contiguous(), s.float().contiguous()))
    return parts
