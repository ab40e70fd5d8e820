"""Per-call timing of the slide pipeline (detect + append + merge) on one GPU."""
import sys, time
import torch
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
from hd_yolo_b200 import ops, synth
from hd_yolo_b200.pipeline import SlidePostprocessor

S = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
dev = torch.device("cuda:0")
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
post = SlidePostprocessor(spec, (S, S), (1024, 1024), 64, 0.25, 0.45, 4096, cap=4096, batch=148, device=dev)
t0, t1 = post.tile_range
store = [synth.slide_tile_logits(post.rois[a:min(a + 148, t1)], 1024, 4, seed=1, first_tile=a, device=dev) for a in range(t0, t1, 148)]
prov = lambda a, b: store[(a - t0) // 148]
post.run(prov)
torch.cuda.synchronize()
for rep in range(2):
    ops.profile.enabled = True
    ops.profile.reset()
    t = time.perf_counter()
    post.detect(prov)
    torch.cuda.synchronize()
    t_det = time.perf_counter() - t
    t = time.perf_counter()
    r = post.merge(ordered=True)
    torch.cuda.synchronize()
    t_mrg = time.perf_counter() - t
    prof = ops.profile.summary()
    ops.profile.enabled = False
print(f"slide {S}: tiles {t1 - t0} rows {r['n']} kept {int((r['state'] == 1).sum())} detect {t_det * 1e3:.1f} ms merge {t_mrg * 1e3:.1f} ms")
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:32s} calls {c:5d} total {ms:10.3f} ms")
