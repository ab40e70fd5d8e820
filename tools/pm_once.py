import sys, torch
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
protos = torch.randn((32, 160, 160), generator=g).to(dev)
k = 50
coef = torch.randn((k, 32), generator=g).to(dev)
c = torch.rand((k, 2), generator=g) * 600 + 20
s = 12 + 24 * torch.rand((k, 2), generator=g)
boxes = torch.cat([c - s / 2, c + s / 2], 1).to(dev)
up = len(sys.argv) > 1 and sys.argv[1] == "up"
m = hdy.process_mask(protos, coef, boxes, (640, 640), upsample=up)
torch.cuda.synchronize()
print("ok", m.shape, float(m.sum()))
