"""Diagnostic for the channels-last tf32x3 path: where does it differ from the exact kernel?"""
import sys, torch
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
bs, C, tile = 2, 128, 320
strides = [8]
feats = [torch.randn((bs, C, tile // 8, tile // 8), generator=g, device=dev)]
K = 64
c = 40 + torch.rand((K, 2), generator=g, device=dev) * 200
sz = 16 + 8 * torch.rand((K, 2), generator=g, device=dev)
img = torch.randint(0, bs, (K, 1), generator=g, device=dev).float()
rois = torch.cat([img, c - sz / 2, c + sz / 2], 1).contiguous()
lv = torch.zeros((K,), device=dev)
ex = hdy.multiscale_roi_align(feats, rois, lv, strides, 14, 2, False)
nc = hdy.multiscale_roi_align(feats, rois, lv, strides, 14, 2, False, mode="tf32x3")
cl = hdy.multiscale_roi_align([f.contiguous(memory_format=torch.channels_last) for f in feats], rois, lv, strides, 14, 2,
                              False, mode="tf32x3")
print("nchw max err", (nc - ex).abs().max().item())
e = (cl - ex).abs()
print("cl max err", e.max().item(), "mean", e.mean().item(), "ref mean abs", ex.abs().mean().item())
print("per 8-channel group max err:", [round(e[:, i:i + 8].max().item(), 3) for i in range(0, C, 8)])
print("per roi max err (first 8):", [round(e[i].max().item(), 3) for i in range(8)])
print("per bin-row max err:", [round(e[:, :, i].max().item(), 3) for i in range(14)])
# does cl match ex under a channel permutation within 32-blocks?  correlate channel j of cl with all channels of ex (roi 0)
a = cl[0].flatten(1); b = ex[0].flatten(1)
corr = torch.corrcoef(torch.cat([a, b]))[:C, C:]
print("best matching exact channel for cl channels 0..15:", corr[:16].argmax(1).tolist())
print("their correlations:", [round(v, 3) for v in corr[:16].max(1).values.tolist()])
print("cl channels 32..47 ->", corr[32:48].argmax(1).tolist())
print("cl channels 64..79 ->", corr[64:80].argmax(1).tolist())
