"""Per-phase SM-cycle profile of hdy_nms_tiles on a synthetic batch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hd_yolo_b200 as hdy
from hd_yolo_b200 import synth, _lib

tile, bs, n_cand, md, cap = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (640, 64, 1000, 1000, 2048)))
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
dets = synth.nuclei_logits(bs, tile, 4, n_cand, seed=1, conf=0.25, generator_device="cuda")
for _ in range(3):
    hdy.detect_postprocess(dets, spec, 0.25, 0.45, md, cap=cap)
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib = _lib.load()
torch.cuda.synchronize()
_lib.check(lib.hdy_debug_nms_phases(_lib.ptr(buf)))
reps = 10
for _ in range(reps):
    out = hdy.detect_postprocess(dets, spec, 0.25, 0.45, md, cap=cap)
torch.cuda.synchronize()
_lib.check(lib.hdy_debug_nms_phases(None))
b = buf.tolist()
ctas = max(b[7], 1)
names = {1: "sort", 3: "binning", 4: "pairs", 5: "resolve", 6: "emit"}
print(f"tile={tile} bs={bs} cand/tile={float(out.cand_counts[:bs].float().mean()):.0f} kept={float(out.counts.float().mean()):.0f}")
for i, nm in names.items():
    print(f"  {nm:8s} {b[i] / ctas:10.0f} cycles/CTA")
print(f"  sweeps/CTA {b[0] / ctas:.2f}   total {sum(b[1:7]) / ctas:.0f} cycles/CTA")
