"""Per-rank, per-call timing of SlidePostprocessor.detect / merge under torchrun (the per-phase table of the sharded
merge in DESIGN.md):   torchrun --nproc-per-node N tools/slide_dist_profile.py [slide_px] [shortcut 0|1]
Every C-ABI call is bracketed by CUDA events (ops.profile); the collectives are timed with events around
comm.all_gather; wall clock around the whole merge shows what the host adds."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hd_yolo_b200 as hdy  # noqa: E402
from hd_yolo_b200 import dist as hdist, ops, synth  # noqa: E402
from hd_yolo_b200.pipeline import SlidePostprocessor  # noqa: E402
from hd_yolo_b200.slide import fold_digest, kept_digest  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
short = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
post = SlidePostprocessor(spec, (S, S), (1024, 1024), 64, 0.25, 0.45, 3328, cap=4096, batch=148, rank=rank, world=world,
                          device=dev, interior_shortcut=short, streams=3)
t0, t1 = post.tile_range
store = [synth.slide_tile_logits(post.rois[a:min(a + 148, t1)], 1024, 4, seed=1, first_tile=a, device=dev)
         for a in range(t0, t1, 148)]
prov = lambda a, b: store[(a - t0) // 148]  # noqa: E731


class TimedComm:
    """all_gather with CUDA events around it."""

    def __init__(self, inner):
        self.inner, self.rank, self.world, self.ev = inner, inner.rank, inner.world, []

    def all_gather(self, t, out=None):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = self.inner.all_gather(t, out)
        b.record()
        self.ev.append((a, b, t.numel() * t.element_size()))
        return r


if world > 1:
    post.comm = TimedComm(post.comm)
for rep in range(5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    post.detect(prov)
    torch.cuda.synchronize()
    td = time.perf_counter() - t
    ops.profile.enabled = rep == 4
    ops.profile.reset()
    if world > 1:
        post.comm.ev.clear()
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = post.merge(ordered=True)
    torch.cuda.synchronize()
    tm = time.perf_counter() - t
    line = {"rank": rank, "rep": rep, "tiles": t1 - t0, "rows": int(r["n"]), "detect_ms": round(td * 1e3, 2),
            "merge_wall_ms": round(tm * 1e3, 2), "seam_rows": r.get("seam_rows"), "exchanges": r.get("exchanges")}
    if rep == 4:
        prof = ops.profile.summary()
        line["calls_ms"] = {k: [n, round(v, 3)] for k, (n, v) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        if world > 1:
            line["all_gathers"] = [[round(a.elapsed_time(b), 3), nb] for a, b, nb in post.comm.ev]
        ops.profile.enabled = False
    print(json.dumps(line), flush=True)
d = kept_digest(r["state"], r["base"])
if world > 1:
    dist.all_reduce(d)
if rank == 0:
    print(json.dumps({"digest": fold_digest(d)}), flush=True)
if world > 1:
    dist.destroy_process_group()
