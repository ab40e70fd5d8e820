import sys, torch, torchvision
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
K, C, bs, tile, strides = 3000, 64, 64, 640, [8, 16, 32]
feats = [torch.randn((bs, C, tile // s, tile // s), generator=g, device=dev) for s in strides]
c = torch.rand((K, 2), generator=g, device=dev) * tile
sz = 12 + 24 * torch.rand((K, 2), generator=g, device=dev)
img = torch.randint(0, bs, (K, 1), generator=g, device=dev).float().sort(0).values
rois = torch.cat([img, c - sz / 2, c + sz / 2], 1).contiguous()
u = torch.rand((K,), generator=g, device=dev)
levels = (u > 0.8).float() + (u > 0.95).float()
ours = hdy.multiscale_roi_align(feats, rois, levels, strides, 14, 2, False)
for i, s in enumerate(strides):
    idx = torch.where(levels == i)[0]
    tv_cuda = torchvision.ops.roi_align(feats[i], rois[idx], (14, 14), 1 / s, 2, False)
    tv_cpu = torchvision.ops.roi_align(feats[i].cpu(), rois[idx].cpu(), (14, 14), 1 / s, 2, False)
    o = ours[idx]
    d1 = (o.cpu() - tv_cpu).abs()
    d2 = (tv_cuda.cpu() - tv_cpu).abs()
    print(f"level {i}: n={len(idx)} ours-vs-cpu max {d1.max().item():.3e} (equal {torch.equal(o.cpu(), tv_cpu)})  cuda-vs-cpu max {d2.max().item():.3e}")
    if d2.max() > 1e-3:
        j = int(d2.view(len(idx), -1).max(1).values.argmax())
        print("   worst roi", rois[idx][j].tolist(), "cuda-vs-cpu per-roi max", d2[j].max().item(), "ours-vs-cpu", d1[j].max().item())
