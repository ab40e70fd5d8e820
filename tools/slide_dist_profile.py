"""Per-rank timing of SlidePostprocessor.detect / merge under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import hd_yolo_b200 as hdy
from hd_yolo_b200 import synth
from hd_yolo_b200.pipeline import SlidePostprocessor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
short = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
post = SlidePostprocessor(spec, (S, S), (1024, 1024), 64, 0.25, 0.45, 4096, cap=4096, batch=128, rank=rank, world=world, device=dev, interior_shortcut=short)
t0, t1 = post.tile_range
store = [synth.slide_tile_logits(post.rois[a:min(a + 128, t1)], 1024, 4, seed=1, first_tile=a, device=dev) for a in range(t0, t1, 128)]
prov = lambda a, b: store[(a - t0) // 128]
for rep in range(4):
    dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    post.detect(prov); torch.cuda.synchronize(); td = time.perf_counter() - t
    t = time.perf_counter(); r = post.merge(ordered=True); torch.cuda.synchronize(); tm = time.perf_counter() - t
    print(f"rank {rank} rep {rep} shortcut {post.shortcut} tiles {t1 - t0} rows {r['n']} detect {td * 1e3:.1f} ms merge {tm * 1e3:.1f} ms", flush=True)
dist.destroy_process_group()
