"""Times hdy.multiscale_roi_align at the tiles640 scale (64 tiles, ~900 detections per tile, 256 channels, 14x14):
the exact-order kernel, the tensor-core form (mode="tf32x3") on NCHW and on channels-last features, and torchvision's own
CUDA roi_align called the way the reference does (one call per level + scatter, yolo_head.py:279-299).  CUDA events,
warm-up, outputs (11.6 GB) far larger than L2.  Usage: python tools/roi_bench.py [K] [C] [bs]; bench.py imports
run_roi() for its `roi_align` sub-record."""
import json
import sys

import torch


def run_roi(K=57600, C=256, bs=64, dev=None, with_torchvision=True):
    import torchvision
    import hd_yolo_b200 as hdy
    dev = dev if dev is not None else torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    tile, strides = 640, [8, 16, 32]
    feats = [torch.randn((bs, C, tile // s, tile // s), generator=g, device=dev) for s in strides]
    feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    c = torch.rand((K, 2), generator=g, device=dev) * tile
    sz = 12 + 24 * torch.rand((K, 2), generator=g, device=dev)               # nuclei: 12-36 px
    img = torch.randint(0, bs, (K, 1), generator=g, device=dev).float().sort(0).values
    rois = torch.cat([img, c - sz / 2, c + sz / 2], 1).contiguous()
    u = torch.rand((K,), generator=g, device=dev)
    levels = (u > 0.8).float() + (u > 0.95).float()                            # 80 / 15 / 5 % on levels 0 / 1 / 2

    def ours(mode="exact", f=feats):
        return hdy.multiscale_roi_align(f, rois, levels, strides, 14, 2, False, mode=mode)

    def reference():
        result = torch.zeros((K, C, 14, 14), device=dev)
        for i, s in enumerate(strides):
            idx = torch.where(levels == i)[0]
            result[idx] = torchvision.ops.roi_align(feats[i], rois[idx], (14, 14), 1 / s, 2, False)
        return result

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    t_ours = timed(ours)
    t_tc = timed(lambda: ours("tf32x3"))
    t_tc_cl = timed(lambda: ours("tf32x3", feats_cl))
    t_ref = timed(reference, 3) if with_torchvision else None
    # agreement on a slice: at the full size torchvision's CUDA kernel indexes its output with a 32-bit int, and level 0
    # alone holds 46 000 x 256 x 196 = 2.3e9 elements -- its result is not usable as a reference there (ours indexes
    # with size_t and is bit-identical to torchvision's CPU op, tests/test_gpu_next.py)
    n_chk = min(K, 4096)
    rc, lc = rois[:n_chk].contiguous(), levels[:n_chk].contiguous()
    o = hdy.multiscale_roi_align(feats, rc, lc, strides, 14, 2, False)
    diff = None
    if with_torchvision:
        r = torch.zeros_like(o)
        for i, s in enumerate(strides):
            idx = torch.where(lc == i)[0]
            r[idx] = torchvision.ops.roi_align(feats[i], rc[idx], (14, 14), 1 / s, 2, False)
        diff = (o - r).abs().max().item()
    o_tc = hdy.multiscale_roi_align(feats, rc, lc, strides, 14, 2, False, mode="tf32x3")
    o_cl = hdy.multiscale_roi_align(feats_cl, rc, lc, strides, 14, 2, False, mode="tf32x3")
    mag = hdy.multiscale_roi_align([f.abs() for f in feats], rc, lc, strides, 14, 2, False)   # sum |w| |f| per output
    out_bytes = K * C * 14 * 14 * 4
    return {"K": K, "C": C, "bs": bs, "out_GB": out_bytes / 1e9,
            "ms_exact": t_ours, "write_GBs_exact": out_bytes / t_ours / 1e6,
            "ms_tf32x3": t_tc, "write_GBs_tf32x3": out_bytes / t_tc / 1e6,
            "ms_tf32x3_channels_last": t_tc_cl, "write_GBs_tf32x3_channels_last": out_bytes / t_tc_cl / 1e6,
            "tf32x3_max_err_over_tap_magnitude_first_4096_rois": ((o_tc - o).abs() / mag.clamp_min(1e-30)).max().item(),
            "tf32x3_channels_last_max_err_over_tap_magnitude_first_4096_rois":
                ((o_cl - o).abs() / mag.clamp_min(1e-30)).max().item(),
            "rows_left_to_the_exact_kernel_first_4096_rois": int((o_tc == o).flatten(1).all(1).sum()),
            "ms_torchvision_cuda_per_level_loop": t_ref,
            "write_GBs_torchvision": (out_bytes / t_ref / 1e6) if t_ref else None,
            "exact_max_abs_diff_vs_torchvision_cuda_first_4096_rois": diff}


if __name__ == "__main__":
    sys.path.insert(0, ".")
    a = [int(v) for v in sys.argv[1:4]]
    print(json.dumps(run_roi(*a)))
