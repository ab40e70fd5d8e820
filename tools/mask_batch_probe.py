"""One slide batch (148 tiles of the synthetic nuclei field) through detect + process_mask_packed, timed per call with
CUDA events; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_yolo_b200 as hdy
from hd_yolo_b200 import synth, masks as hm
from hd_yolo_b200.slide import sliding_window_scanner

dev = torch.device("cuda:0")
bs, tile, nm = 148, 1024, 32
md = int(sys.argv[1]) if len(sys.argv) > 1 else 3328
rois = sliding_window_scanner((100000, 100000), (tile, tile), 64)[1000:1000 + bs]
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4, no=9 + nm)
dets = synth.slide_tile_logits(rois, tile, 4, seed=1, first_tile=1000, extra=nm, device=dev)
protos = synth.slide_tile_protos(bs, tile, seed=1, first_tile=1000, device=dev)
out = hdy.detect_postprocess(dets, spec, 0.25, 0.45, md, cap=4096)
print("kept/tile", float(out.counts.float().mean()), "max", int(out.counts.max()))
for rep in range(3):
    pm = hm.process_mask_packed(protos, out.extra, out.boxes, out.counts, (tile, tile), upsample=True,
                                capacity_words=int(bs * md * 40))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for rep in range(5):
    pm = hm.process_mask_packed(protos, out.extra, out.boxes, out.counts, (tile, tile), upsample=True,
                                capacity_words=int(bs * md * 40))
e1.record()
torch.cuda.synchronize()
pm.check()
print("process_mask_packed (geometry + masks) ms per batch:", e0.elapsed_time(e1) / 5, "words", int(pm.offsets[-1]))
# how many detections the region path handed to the per-detection kernel (first word of the workspace), and their sizes
ws, _ = hm._pm_workspace(dev, bs, md)
n_listed = int(ws.view(torch.int32)[0])
w = (out.boxes[..., 2] - out.boxes[..., 0])
h = (out.boxes[..., 3] - out.boxes[..., 1])
valid = torch.arange(md, device=dev)[None, :] < out.counts[:, None]
big = valid & ((w > 64) | (h > 64))
print("listed detections in the batch:", n_listed, "| boxes over 64 px:", int(big.sum()),
      "| largest box side:", float(torch.maximum(w, h)[valid].max()),
      "| sizes of the big ones:", torch.maximum(w, h)[big].sort(descending=True).values[:12].tolist())
