"""CUDA-event timing of the slide merge in steps (build / each round / finish) on the real slide field."""
import sys
import torch
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
from hd_yolo_b200 import dist as hdist, synth
from hd_yolo_b200.pipeline import SlidePostprocessor
from hd_yolo_b200.slide import dirty_tiles, tile_cores

S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = torch.device("cuda:0")
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
post = SlidePostprocessor(spec, (S, S), (1024, 1024), 64, 0.25, 0.45, 4096, cap=4096, batch=148, device=dev)
t0, t1 = post.tile_range
store = [synth.slide_tile_logits(post.rois[a:min(a + 148, t1)], 1024, 4, seed=1, first_tile=a, device=dev) for a in range(t0, t1, 148)]
post.detect(lambda a, b: store[(a - t0) // 148])
acc = post.acc
n = acc.count()
boxes, scores = acc.boxes[:n], acc.scores[:n]
rois = torch.cat(acc.rois)
cores = tile_cores(rois).to(dev)
margin, fb, ft, fc = acc.overhang()
dirty = dirty_tiles(fb, ft, fc, rois)
gidx = torch.arange(n, device=dev, dtype=torch.int32)


def ev():
    return torch.cuda.Event(enable_timing=True)


for rep in range(3):
    e = [ev() for _ in range(12)]
    e[0].record()
    be = hdist.DeviceMergeBackend(boxes, scores, gidx, n, 0.25, 0.45, tile_id=acc.tile[:n], cores=cores, margin=margin,
                                  dirty=dirty)
    e[1].record()
    for r in range(8):
        be.rounds(r, 1)
        e[2 + r].record()
    st, ok = be.finish()
    e[10].record()
    torch.cuda.synchronize()
    ts = [e[i].elapsed_time(e[i + 1]) for i in range(10)]
active = int((be.state[:n] == 0).sum()) if False else None
print(f"n={n} margin={float(margin):.2f} far={int(fc)} build {ts[0]:.3f} rounds {[round(x, 3) for x in ts[1:9]]} finish {ts[9]:.3f} "
      f"ok={ok} kept={int((st == 1).sum())} suppressed={int((st == 2).sum())}")
