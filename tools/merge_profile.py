"""Times hdy_merge_build / each round / finish on synthetic banded detections."""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from hd_yolo_b200 import dist as hdist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
side = int((n ** 0.5) * 18.7)
c = torch.rand((n, 2), generator=g, device=dev) * side
sz = 12 + 18 * torch.rand((n, 2), generator=g, device=dev)
boxes = torch.cat([c - sz / 2, c + sz / 2], 1).contiguous()
scores = 0.3 + 0.6 * torch.rand((n,), generator=g, device=dev)
gidx = torch.arange(n, device=dev, dtype=torch.int32)
def T(f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r
for rep in range(2):
    tb, be = T(lambda: hdist.DeviceMergeBackend(boxes, scores, gidx, n, 0.25, 0.45))
    tr = [T(lambda r=r: be.rounds(r, 1))[0] for r in range(8)]
    tf, (st, ok) = T(be.finish)
print(f"n={n} side={side} build {tb:.2f} ms rounds {[round(x, 2) for x in tr]} finish {tf:.2f} ms ok={ok} kept={int((st == 1).sum())}")
