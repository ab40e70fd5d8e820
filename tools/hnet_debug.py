import sys, torch
sys.path.insert(0, ".")
from hd_yolo_b200 import hnet, synth_hnet as sh
from oracle import port
dev = torch.device("cuda:0")
n_img, size, pre, post = 2, 256, 1000, 1000
anchors, counts, obj, deltas = sh.rpn_inputs(n_img, size, seed=size + n_img)
shapes = [(size, size - 16)] * n_img
prop_ref = port.rcnn_box_decode(deltas, [anchors] * n_img).view(n_img, -1, 4)
rb, rs = port.rpn_filter_proposals(prop_ref, obj, shapes, counts, pre, post, 0.7, 0.0)
for mode in ("torchvision-cpu", "vanilla"):
    gb, gs = hnet.rpn_filter_proposals(prop_ref.to(dev), obj.to(dev), shapes, counts, pre, post, 0.7, 0.0, mode=mode)
    for i in range(n_img):
        a, b = gb[i].cpu(), rb[i]
        bad = (a != b).any(1)
        print(mode, i, a.shape, b.shape, "rows differing:", int(bad.sum()), "first:", torch.nonzero(bad).flatten()[:10].tolist())
        j = torch.nonzero(bad).flatten()[:4]
        print("  ours", a[j].tolist(), gs[i].cpu()[j].tolist())
        print("  ref ", b[j].tolist(), rs[i][j].tolist())
        sa = set(map(tuple, a.tolist())); sb = set(map(tuple, b.tolist()))
        print("  set diff", len(sa - sb), len(sb - sa))
