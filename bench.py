#!/usr/bin/env python
"""Benchmark of the post-processing hot path (BASELINE.json metric: post-proc tiles/s & boxes/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tiles640|tiles1024|slide] [--masks proto|none]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

tiles640 / tiles1024 (BASELINE.json configs[1] / configs[2]): one "step" = one pass of the hot path over one batch of
synthetic head outputs: fused decode+filter+compact -> per-tile NMS -> score/label select -> proto masks
(32-prototype contraction, sigmoid, crop, bilinear upsample, threshold, bit-packed).  Weak scaling over GPUs (every
rank processes its own batches; the path has no cross-tile dependency).

slide (configs[3]): one "step" = a whole synthetic 100k x 100k px slide (11 025 tiles of 1024 px, 64 px overlap) cut
from one global nuclei field: per-tile post-processing of this rank's tile rows, append in slide coordinates, exact
slide-level merge NMS with the seam exchange over NCCL.  Strong scaling.  Unless --no-slide, a short slide run is also
attached to the tiles line as "slide".

Prints ONE JSON line on rank 0:
  value     : whole-job tiles/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       : the same through the public API with HOST (pinned) inputs: H2D of the step's inputs and D2H of its
              results inside the timed region
  roofline  : the dominant HBM-bound kernel: algorithmic bytes per launch / its CUDA-event time (measured live)
  stages    : every C-ABI call of the step with its time, algorithmic bytes and achieved GB/s
  cpu_baseline : oracle port (the reference's torch/torchvision CPU path) on a bounded sample
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

# ... and when the environment asks for NCCL_DEBUG=VERSION/INFO anyway (or any library writes to fd 1): the JSON line
# goes to a private duplicate of stdout, everything else that is written to fd 1 lands on stderr.
_JSON_FD = None


def _protect_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)

WORKLOADS = {
    # BASELINE.json configs[1]: 640x640 tiles, batch 64, 3 anchor levels, 32 prototypes, ~1k candidates/tile
    "tiles640": dict(tile=640, bs=64, n_cand=1000, conf=0.25, iou=0.45, max_det=1000, nc=4, cap=2048),
    # BASELINE.json configs[2]: 1024x1024 dense-nuclei tiles, batch 128, ~3k candidates/tile
    "tiles1024": dict(tile=1024, bs=128, n_cand=3000, conf=0.25, iou=0.45, max_det=3000, nc=4, cap=4096),
    # BASELINE.json configs[3]: whole slide, 1024-px tiles, 64-px overlap, ~3k candidates/tile
    # (148 tiles per batch: the per-tile NMS runs one CTA per tile, one per SM)
    "slide": dict(tile=1024, bs=148, n_cand=3000, conf=0.25, iou=0.45, max_det=4096, nc=4, cap=4096, overlap=64),
}
NM = 32            # prototypes
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiles640", choices=list(WORKLOADS))
    ap.add_argument("--masks", default="proto", choices=["proto", "paste", "none"])
    ap.add_argument("--slide-size", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slide", action="store_true")
    ap.add_argument("--layout", type=int, default=0, choices=[0, 1],
                    help="tiles workloads: 0 = the reference's permuted [bs,na,ny,nx,no] head tensors, 1 = the 1x1 "
                         "conv's native [bs,na*no,ny,nx] output (the yolo_head.py:141-145 permute copy skipped)")
    ap.add_argument("--slide-streams", type=int, default=3,
                    help="slide workload: tile batches alternate over this many streams (SlidePostprocessor(streams=))")
    ap.add_argument("--inflight", type=int, default=3, help="steps in flight (CUDA graphs on separate streams)")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 5 if a.workload == "slide" else 200
    if a.warmup is None:
        a.warmup = 3 if a.workload == "slide" else 10
    return a


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed regions run (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_pass(dets_cpu, protos_cpu, wl, spec_args, masks):
    """The reference's CPU path for one batch: compute_proposals -> pad/cat -> nms_per_image -> select
    (-> process_mask with upsample per tile)."""
    from oracle import port
    anchors, strides = spec_args
    nc = wl["nc"]
    preds = port.compute_proposals(dets_cpu, anchors, strides)
    params = {'conf_thres': wl["conf"], 'iou_thres': wl["iou"], 'max_det': wl["max_det"]}
    if masks == "none":
        return port.compute_outputs(preds, nc, params)
    import torch
    cat = port.concat_levels(preds)
    res = port.nms_per_image(cat, nc, wl["conf"], wl["iou"], wl["max_det"])
    out = []
    for i, r in enumerate(res):
        s, l = port.select_scores(r['scores'][:, :1 + nc].clone(), wl["conf"], port.default_descendants(nc))
        if masks == "proto":
            coef = r['extra'][:, :NM]      # extra = raw coefficient channels + level id; the coefficients come first
            m = port.process_mask(protos_cpu[i], coef, r['boxes'], (wl["tile"], wl["tile"]), upsample=True)
        else:                              # reference variant A: mask select + paste_masks_in_image, > 0.5
            k = len(r['boxes'])
            sel = port.mask_select(protos_cpu[i][:k], l, torch.tensor([-1, 0, 0, 1, 1]))
            m = port.paste_masks_in_image(sel, r['boxes'], (wl["tile"], wl["tile"]), padding=1) > 0.5
        out.append((r['boxes'], s, l, m))
    return out


def cpu_sample_inputs(wl, masks, sample, seed=1):
    import torch
    from hd_yolo_b200 import synth
    extra = NM if masks == "proto" else 0
    dets = synth.nuclei_logits(sample, wl["tile"], wl["nc"], wl["n_cand"], seed=seed, conf=wl["conf"], extra=extra)
    protos = None
    if masks == "proto":
        g = torch.Generator().manual_seed(seed + 7)
        protos = torch.randn((sample, NM, wl["tile"] // 4, wl["tile"] // 4), generator=g)
    elif masks == "paste":
        g = torch.Generator().manual_seed(seed + 8)
        protos = torch.randn((sample, min(wl["max_det"], wl["cap"]), 2, 28, 28), generator=g)
    return dets, protos


def time_cpu(wl, masks, budget_s, max_reps, threads=None):
    """Bounded CPU timing of the oracle port on `threads` host threads (default: all).  Returns (tiles/s, threads,
    sample text)."""
    import torch
    from hd_yolo_b200 import synth
    torch.set_num_threads(threads or os.cpu_count() or 1)
    sample = 8 if masks == "none" else 2
    dets, protos = cpu_sample_inputs(wl, masks, sample)
    spec_args = (synth.ANCHORS_3, synth.STRIDES_3)
    cpu_reference_pass(dets, protos, wl, spec_args, masks)          # warm-up
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        cpu_reference_pass(dets, protos, wl, spec_args, masks)
        reps += 1
    dt = time.perf_counter() - t0
    what = "compute_proposals + nms_per_image + score select" + \
        {"proto": " + process_mask(upsample)", "paste": " + mask_select + paste_masks_in_image", "none": ""}[masks]
    return sample * reps / dt, torch.get_num_threads(), (
        f"{sample} tiles x {reps} passes of oracle/port.py ({what}; torch {torch.__version__} CPU + torchvision nms)")


def cpu_merge_scaling(boxes, scores, tile, n_cols, conf, iou, grids=(2, 3, 4, 6, 8), budget_s=25.0):
    """the reference's slide merge (Ensemble.merge, yolo.py:165-204 = one dense torchvision.ops.nms over
    everything merge_outputs concatenated) is O(n^2) -- 338 s at 2e5 boxes -- so it is timed on g x g tile sub-slides
    of the SAME detections and reported with its measured exponent instead of an extrapolated full-slide number.
    boxes [n,4] / scores [n] / tile [n] (tile index in the slide's row-major grid of n_cols columns): CPU tensors in
    merge_outputs' order.  Returns a dict for the JSON line."""
    import math
    import torch
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    rows, cols = tile // n_cols, tile % n_cols
    pts, t_start = [], time.perf_counter()
    for g in grids:
        sel = (rows < g) & (cols < g)
        b, sc = boxes[sel].contiguous(), scores[sel].contiguous()
        if len(b) < 16:
            continue
        if pts and time.perf_counter() - t_start + pts[-1][2] * (len(b) / pts[-1][1]) ** 2 > budget_s:
            break                                   # the next size would not fit the budget
        labels = torch.zeros((len(b),), dtype=torch.int64)
        t0 = time.perf_counter()
        r = port.ensemble_merge([{'det': {'boxes': b, 'scores': sc, 'labels': labels}}],
                                {'conf_thres': conf, 'iou_thres': iou, 'max_det': 10 ** 9})['det']
        pts.append((g * g, int(len(b)), time.perf_counter() - t0, int(len(r['boxes']))))
    out = {"what": "oracle port of Ensemble.merge (dense torchvision.ops.nms) on g x g tile sub-slides of the bench's "
                   "own detections, all host threads", "cores": torch.get_num_threads(),
           "tiles": [p[0] for p in pts], "boxes": [p[1] for p in pts], "seconds": [round(p[2], 4) for p in pts],
           "kept": [p[3] for p in pts]}
    if len(pts) >= 2:
        xs = [math.log(p[1]) for p in pts]
        ys = [math.log(max(p[2], 1e-9)) for p in pts]
        mx, my = sum(xs) / len(xs), sum(ys) / len(ys)
        den = sum((x - mx) ** 2 for x in xs)
        out["exponent"] = round(sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / den, 3) if den > 0 else None
    return out


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation (oracle port: same torch/torchvision calls) on all
    host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    masks = args.masks if args.workload != "slide" else "none"
    v, cores, sample = time_cpu(wl, masks, budget_s=20.0, max_reps=max(2, min(args.steps, 50)))
    line = {
        "impl": "reference", "metric": "postproc_tiles_per_s", "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wl["bs"] / v, "higher_is_better": True,
        "scaling": "strong" if args.workload == "slide" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "tile": wl["tile"], "candidates_per_tile": wl["n_cand"],
                   "conf": wl["conf"], "iou": wl["iou"], "max_det": wl["max_det"], "masks": masks,
                   "note": "ms_per_step is the CPU time for one full batch of the workload, extrapolated from the sample"},
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "boxes_per_s": v * wl["n_cand"],
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------ helpers
class Ctx:
    pass


try:
    _ORIG_AFFINITY = os.sched_getaffinity(0)
except Exception:   # not on Linux
    _ORIG_AFFINITY = set()


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs next to GPU `index` (sysfs local_cpulist of its PCI function) before any pinned
    host buffer is allocated: first-touch then places the e2e staging buffers on the GPU's own NUMA node, which is
    worth up to 1.4x of H2D bandwidth on a two-socket box.  Returns a short description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return f"bound to the {len(cpus)} CPUs local to GPU {index}"
        return "all CPUs are local to the GPU (single NUMA node)"
    except Exception as e:  # no sysfs entry / no PCI ids: leave the affinity alone
        return f"unchanged ({type(e).__name__})"


def make_ctx(args):
    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(c.local)
    c.affinity = bind_to_gpu_numa_node(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    assert c.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={c.world}: launch with torch.distributed.run"

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, side_streams=()):
        """CUDA events around `steps` calls, barrier + synchronize on both sides, max over ranks (ms).  Work that fn
        issues on side_streams is ordered after the start event and before the end event."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record()
        for s_ in side_streams:
            s_.wait_event(e0)
        for i in range(steps):
            fn(i)
        for s_ in side_streams:
            main.wait_stream(s_)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if c.world > 1:
            t = torch.tensor([ms], device=c.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    c.barrier, c.timed = barrier, timed
    return c


def load_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (of fallback)"


# ------------------------------------------------------------------------------------------------ tiles workloads
def run_tiles(args, wl, c):
    import torch
    import hd_yolo_b200 as hdy
    from hd_yolo_b200 import masks as hmasks
    from hd_yolo_b200 import ops, synth

    masks = args.masks
    extra = NM if masks == "proto" else 0
    nc, tile, bs, K, W = wl["nc"], wl["tile"], wl["bs"], args.steps, max(args.warmup, 3)
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=nc, no=5 + nc + extra)
    shapes = synth.level_shapes(tile, synth.STRIDES_3)
    N = spec.rows_per_tile(shapes)
    mh = mw = tile // 4
    in_bytes = bs * N * spec.no * 4 + (bs * NM * mh * mw * 4 if masks == "proto" else 0) + \
        (bs * min(wl["max_det"], wl["cap"]) * 2 * 28 * 28 * 4 if masks == "paste" else 0)
    R = max(2, args.inflight, int(2.5 * L2_BYTES / in_bytes) + 1)   # rotate input batches so that reads miss L2
    batches = [synth.nuclei_logits(bs, tile, nc, wl["n_cand"], seed=1000 * c.rank + r, conf=wl["conf"], extra=extra,
                                   generator_device="cuda") for r in range(R)]
    if args.layout == 1:   # the same logits as the head's 1x1 conv leaves them: [bs, na*no, ny, nx]
        batches = [[d.permute(0, 1, 4, 2, 3).reshape(d.shape[0], -1, d.shape[2], d.shape[3]).contiguous() for d in b]
                   for b in batches]
    protos = None
    if masks == "proto":
        g = torch.Generator(device="cuda").manual_seed(77 + c.rank)
        protos = [torch.randn((bs, NM, mh, mw), generator=g, device=c.dev) for _ in range(R)]
    md_slots = min(wl["max_det"], wl["cap"])
    mlogits = None
    if masks == "paste":
        # reference variant A (yolo_head.py:321-353 + torchvision paste_masks_in_image): one 2-channel 28x28 mask
        # logit map per detection slot, as the RoI mask head would produce them
        g = torch.Generator(device="cuda").manual_seed(78 + c.rank)
        mlogits = [torch.randn((bs * md_slots, 2, 28, 28), generator=g, device=c.dev) for _ in range(R)]
        mask_indices = torch.tensor([-1, 0, 0, 1, 1], dtype=torch.int32, device=c.dev)
        slot_ids = torch.arange(md_slots, device=c.dev)[None, :]
    state = {"cap_words": None}

    def step(i, dets=None, pr=None):
        out = hdy.detect_postprocess(dets if dets is not None else batches[i % R], spec, wl["conf"], wl["iou"],
                                     wl["max_det"], cap=wl["cap"], layout=args.layout)
        pm = None
        if masks == "proto":
            pm = hmasks.process_mask_packed(pr if pr is not None else protos[i % R], out.extra, out.boxes, out.counts,
                                            (tile, tile), upsample=True, capacity_words=state["cap_words"])
        elif masks == "paste":
            # M1: channel = mask_indices[labels.clamp(min=0)], empty beyond counts; M2+M3 fused and bit-packed
            live = slot_ids < out.counts[:, None]            # slots beyond counts hold uninitialised labels
            ch = mask_indices[torch.where(live, out.labels, torch.zeros_like(out.labels)).clamp(min=0)]
            ch = torch.where(live, ch, torch.full_like(ch, -1)).reshape(-1)
            pm = hmasks.paste_masks_packed(pr if pr is not None else mlogits[i % R], out.boxes.reshape(-1, 4),
                                           (tile, tile), padding=1, channel=ch, apply_sigmoid=True,
                                           capacity_words=state["cap_words"])
        return out, pm

    # size the packed-mask buffer once (the only call that reads a size back), then never sync inside a step
    out, pm = step(0)
    out.to_list()                      # raises on candidate-capacity overflow
    words = 0
    if pm is not None:
        words = int(pm.offsets[-1].item())
        state["cap_words"] = int(words * 1.2) + 1024

    sampler = ClockSampler(c.local)
    sampler.start()
    for i in range(W):
        out, pm = step(i)
    if pm is not None:
        pm.check()
    # The step is launch-bound on the host at these sizes (7 C-ABI calls + ~25 small allocations per step), so the
    # public API offers CUDA-graph capture of a fixed-shape step (hdy.CapturedStep); one graph per rotating input batch.
    # Every graph has its own scratch slot, so two steps can be in flight on two streams: the per-tile NMS occupies
    # only `bs` of the 148 SMs and the tails of the other kernels leave SMs idle, which the neighbouring step fills.
    graphs = [hdy.CapturedStep(lambda r=r: step(r), slot=r) for r in range(R)]
    side = [torch.cuda.Stream() for _ in range(max(1, args.inflight))]

    def gstep(i):
        return graphs[i % R]()

    def gstep2(i):
        return graphs[i % R](stream=side[(i % R) % len(side)])

    for i in range(W):
        gstep(i)
    ms_serial = c.timed(gstep, K)
    for i in range(W):
        gstep2(i)
    torch.cuda.synchronize()
    ops.profile.reset()
    ms = c.timed(gstep2, K, side_streams=side)
    launches = ops.profile.launches
    tiles_per_s = c.world * bs * K / (ms * 1e-3)
    torch.cuda.synchronize()
    out, pm = gstep(0)

    # ---- per-call CUDA-event times of the same step, un-graphed.  A GPU-side sleep is queued first so that the
    #      host runs ahead and the event pairs bracket back-to-back kernels, not launch gaps.
    ops.profile.enabled = True
    ops.profile.reset()

    def pstep(i):
        torch.cuda._sleep(3_000_000)
        step(i)

    c.timed(pstep, min(K, 50))
    prof = ops.profile.summary()
    ops.profile.enabled = False
    cand = float(out.cand_counts[:bs].float().mean())
    kept = float(out.counts.float().mean())
    ne = spec.no - 5 - nc
    alg = {   # algorithmic bytes per launch
        # layout 1: only the objectness plane is streamed, survivors gather their four box logits
        "hdy_filter_compact_logits": bs * (4 * N * spec.no + 24 * cand) if args.layout == 0
        else bs * (4 * N + (16 + 24) * cand),
        "hdy_nms_tiles": bs * (24 * cand + 28 * kept),
        "hdy_gather_logits": bs * kept * (4 * (1 + nc + ne) * 2 + 4),
        "hdy_gather_select_logits": bs * kept * (4 * (1 + nc + ne) * 2 + 4 + 4 + 8),
        "hdy_select_scores": bs * kept * (4 * (1 + nc) * 2 + 4 + 8),
        "hdy_process_mask_geometry": bs * min(wl["max_det"], wl["cap"]) * (16 + 16 + 8),
        "hdy_process_mask_packed": bs * (4 * NM * mh * mw + kept * (4 * NM + 16 + 8)) + 4 * words,
        "hdy_zero_i32": 4 * (bs + 1),
        "hdy_paste_geometry": bs * md_slots * (16 + 16 + 8 + 4),
        "hdy_paste_masks_packed": bs * kept * (28 * 28 * 4 + 16 + 8) + 4 * words,
    }
    peak, peak_src = load_peak()
    stages = {}
    for name, (calls, tot) in prof.items():
        t = tot / calls
        a = alg.get(name)
        stages[name] = {"ms": t, "alg_bytes": a, "gbs": (a / (t * 1e-3) / 1e9) if a else None}
    hbm_bound = [k for k in stages if k in ("hdy_filter_compact_logits", "hdy_process_mask_packed",
                                            "hdy_paste_masks_packed")]
    dom = max(hbm_bound, key=lambda k: stages[k]["ms"])
    dom_ms, dom_bytes = stages[dom]["ms"], alg[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    step_bytes = sum(v["alg_bytes"] for v in stages.values() if v["alg_bytes"])
    traffic = None
    try:   # DRAM bytes per launch of the dominant call, from the committed ncu --set full capture of this workload
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[args.workload][dom]["bytes"]
        if masks != "proto" and dom == "hdy_filter_compact_logits":
            traffic = None if args.workload != "tiles1024" else traffic
    except Exception:
        pass

    # ---- e2e: host (pinned) inputs, H2D + D2H inside the timed region ---------------------------------------
    e2e = None
    if not args.no_e2e:
        host = [[d.cpu().pin_memory() for d in b] for b in batches[:2]]
        second = protos if protos is not None else mlogits
        host_p = [p.cpu().pin_memory() for p in second[:2]] if second is not None else None
        stage = [torch.empty_like(d) for d in batches[0]]
        stage_p = torch.empty_like(second[0]) if second is not None else None
        md = out.max_det
        h = {"boxes": torch.empty((bs, md, 4), dtype=torch.float32).pin_memory(),
             "scores": torch.empty((bs, md), dtype=torch.float32).pin_memory(),
             "labels": torch.empty((bs, md), dtype=torch.int64).pin_memory(),
             "counts": torch.empty((bs,), dtype=torch.int32).pin_memory()}
        if pm is not None:
            h["geom"] = torch.empty(tuple(pm.geom.shape), dtype=torch.int32).pin_memory()
            h["offsets"] = torch.empty(tuple(pm.offsets.shape), dtype=torch.int64).pin_memory()
            h["bits"] = torch.empty((state["cap_words"],), dtype=torch.int32).pin_memory()

        # Two staging sets and two graphs: the H2D copy of step i+1 runs on a copy stream while step i computes and
        # reads back (PCIe is full duplex), as a serving loop would; every byte of every step is still copied inside
        # the timed region.
        stage2 = [torch.empty_like(d) for d in batches[0]]
        stage_p2 = torch.empty_like(second[0]) if second is not None else None
        stages_e = [(stage, stage_p), (stage2, stage_p2)]
        graphs_e = [hdy.CapturedStep(lambda: step(0, stage, stage_p), slot=0),
                    hdy.CapturedStep(lambda: step(0, stage2, stage_p2), slot=1)]
        copy_stream = torch.cuda.Stream()
        ready = [None, None]     # H2D of the set finished
        freed = [None, None]     # the graph that read the set finished

        def issue_h2d(i):
            k = i % 2
            st, stp = stages_e[k]
            with torch.cuda.stream(copy_stream):
                if freed[k] is not None:
                    copy_stream.wait_event(freed[k])
                for s, hh in zip(st, host[i % 2]):
                    s.copy_(hh, non_blocking=True)
                if stp is not None:
                    stp.copy_(host_p[i % 2], non_blocking=True)
                ready[k] = torch.cuda.Event()
                ready[k].record(copy_stream)

        pending = {"next": None}

        def e2e_step(i):
            if pending["next"] != i:      # first step of a run: nothing was prefetched
                issue_h2d(i)
            k = i % 2
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[k])
            o, p = graphs_e[k]()
            freed[k] = torch.cuda.Event()
            freed[k].record(cur)
            issue_h2d(i + 1)              # overlaps the read-back below and the next replay's wait
            pending["next"] = i + 1
            h["boxes"].copy_(o.boxes, non_blocking=True)
            h["scores"].copy_(o.scores, non_blocking=True)
            h["labels"].copy_(o.labels, non_blocking=True)
            h["counts"].copy_(o.counts, non_blocking=True)
            if p is not None:
                h["geom"].copy_(p.geom, non_blocking=True)
                h["offsets"].copy_(p.offsets, non_blocking=True)
                h["bits"][:p.bits.numel()].copy_(p.bits, non_blocking=True)

        Ke = max(3, min(K, 50))
        for i in range(3):
            e2e_step(i)
        pending["next"] = None
        torch.cuda.synchronize()
        ms_e = c.timed(e2e_step, Ke, side_streams=(copy_stream,))
        d2h = sum(t.numel() * t.element_size() for k, t in h.items() if k != "bits") + (words * 4 if pm is not None else 0)
        e2e = {"value": c.world * bs * Ke / (ms_e * 1e-3), "unit": "tiles/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": ms_e / Ke}
        del host, host_p, stage, stage_p, stage2, stage_p2, stages_e, graphs_e, h
    clocks = sampler.stop()

    line = {
        "metric": "postproc_tiles_per_s", "value": tiles_per_s, "unit": "tiles/s", "n_gpus": c.world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "tile": tile, "tiles_per_step_per_gpu": bs, "levels": 3,
                   "layout": "permuted [bs,na,ny,nx,no] (what the reference's Detect.forward hands over)"
                   if args.layout == 0 else "conv-native [bs,na*no,ny,nx] (yolo_head.py:141-145 permute skipped)",
                   "rows_per_tile": N, "channels": spec.no, "prototypes": NM if masks == "proto" else 0,
                   "candidates_per_tile": round(cand, 1), "kept_per_tile": round(kept, 1), "conf": wl["conf"],
                   "iou": wl["iou"], "max_det": wl["max_det"], "cap": wl["cap"],
                   "stages": "decode+filter+compact, nms, score/label select" +
                             (", process_mask (proto contraction, sigmoid, crop, upsample, >0.5, bit-packed)" if masks == "proto" else "") +
                             (", mask select + paste_masks_in_image > 0.5 (28x28 mask logits, bit-packed)" if masks == "paste" else ""),
                   "l2": f"{R} rotating input batches ({R * in_bytes / 1e6:.0f} MB) > 126 MB L2",
                   "launch": f"one CUDA graph per input batch (hdy.CapturedStep), replayed; {len(side)} steps in "
                             "flight on as many streams (each graph has its own scratch slot)",
                   "ms_per_step_one_stream": ms_serial / K},
        "boxes_per_s": tiles_per_s * cand,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0, "traffic": traffic,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms},
        "stages": stages,
        "pipeline": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms / K * 1e-3) / 1e9,
                     "frac_of_peak": step_bytes / (ms / K * 1e-3) / 1e9 / peak},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    del batches, protos, mlogits, graphs
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------ slide workload
def run_slide(args, wl, c, steps, warmup, want_e2e, want_cpu_merge=False):
    import torch
    import hd_yolo_b200 as hdy
    from hd_yolo_b200 import ops, synth
    from hd_yolo_b200.pipeline import SlidePostprocessor

    nc, tile, bs = wl["nc"], wl["tile"], wl["bs"]
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=nc)
    S = args.slide_size
    post = SlidePostprocessor(spec, (S, S), (tile, tile), wl["overlap"], wl["conf"], wl["iou"], wl["max_det"],
                              cap=wl["cap"], batch=bs, rank=c.rank, world=c.world, device=c.dev,
                              capacity=None, streams=args.slide_streams)
    t0, t1 = post.tile_range
    n_tiles = int(post.rois.shape[0])
    # this rank's head outputs for the whole slide, resident in HBM (N=1: 11 025 tiles x 2.3 MB = 25.6 GB)
    store = []
    for a in range(t0, t1, bs):
        b = min(a + bs, t1)
        store.append(synth.slide_tile_logits(post.rois[a:b], tile, nc, seed=1, first_tile=a, conf=wl["conf"], device=c.dev))
    in_bytes = sum(sum(t.numel() * 4 for t in dets) for dets in store)

    def provider(a, b):
        return store[(a - t0) // bs]

    res = {}

    def step(i):
        res["r"] = post.run(provider, ordered=True)

    for i in range(max(warmup, 4)):     # the slide-sized temporaries settle in the caching allocator after ~3 passes
        step(i)
    ops.profile.reset()
    ms = c.timed(step, steps)
    launches = ops.profile.launches
    # merge alone (detections already appended by the last step)
    ms_merge = c.timed(lambda i: post.merge(ordered=True), max(2, min(steps, 5)))
    ms_merge /= max(2, min(steps, 5))
    r = res["r"]
    n_local = int(r["n"])
    kept_local = int((r["state"] == 1).sum())
    cpu_merge = None
    if want_cpu_merge and c.rank == 0 and c.world == 1:
        try:   # a reported baseline must never take the GPU record down with it
            n_cols = int((post.rois[:, 1] == post.rois[0, 1]).sum())        # tiles in the first grid row
            tl = post.acc.tile[:n_local]
            tl = torch.where(tl >= 0, tl, ~tl).to(torch.int64)              # fragile rows carry ~tile
            sub = ((tl // n_cols) < 8) & ((tl % n_cols) < 8)      # the largest sub-slide cpu_merge_scaling may time
            cpu_merge = cpu_merge_scaling(post.acc.boxes[:n_local][sub].cpu(), post.acc.scores[:n_local][sub].cpu(),
                                          tl[sub].cpu(), n_cols, wl["conf"], wl["iou"])
        except Exception as e:   # noqa: BLE001
            cpu_merge = {"error": f"{type(e).__name__}: {e}"}
    tot = torch.tensor([n_local, kept_local, in_bytes], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot)
    out = {"tiles": n_tiles, "tiles_per_s": n_tiles * steps / (ms * 1e-3), "ms_per_slide": ms / steps,
           "merge_ms": ms_merge, "detections": int(tot[0]), "kept": int(tot[1]),
           "input_bytes": int(tot[2]), "slide_px": S, "tile": tile, "overlap": wl["overlap"],
           "seam_rows": r.get("seam_rows"), "exchanges": r.get("exchanges"), "gpu_launches": launches,
           "boxes_per_s": int(tot[0]) * steps / (ms * 1e-3)}
    if cpu_merge is not None:
        out["cpu_merge"] = cpu_merge
    if want_e2e:
        # host-resident head outputs: every batch is copied H2D inside the timed region (4 pinned host batches are
        # reused in rotation, bytes are counted for every copy), the slide's survivors are read back D2H
        nrot = min(4, len(store))
        host = [[t.cpu().pin_memory() for t in store[k]] for k in range(nrot)]
        # one staging buffer set per stream of the post-processor (batch k runs on stream k % streams)
        stages = [[torch.empty_like(t) for t in store[0]] for _ in range(max(1, args.slide_streams))]

        def provider_h(a, b):
            k = ((a - t0) // bs)
            src = host[k % nrot]
            stage = stages[k % len(stages)]
            n = b - a
            for s, h in zip(stage, src):
                s[:n].copy_(h[:n], non_blocking=True)
            return [s[:n] for s in stage] if n < bs else stage

        hb, pinned = {}, {}

        def e2e_step(i):
            # survivors are read back into pinned host buffers that persist across slides (a pageable .cpu() of
            # 0.8 GB costs more than the whole H2D stream: fresh pages + a staged copy)
            rr = post.run(provider_h, ordered=True)
            for k in ("boxes", "scores", "labels"):
                t = rr[k]
                buf = pinned.get(k)
                if buf is None or buf.shape[0] < t.shape[0]:
                    buf = torch.empty((int(t.shape[0] * 1.05) + 1,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory()
                    pinned[k] = buf
                buf[:t.shape[0]].copy_(t, non_blocking=True)
                hb[k] = buf[:t.shape[0]]

        e2e_step(0)
        Ke = max(1, min(steps, 3))
        ms_e = c.timed(e2e_step, Ke)
        d2h = sum(t.numel() * t.element_size() for t in hb.values())
        out["e2e"] = {"value": n_tiles * Ke / (ms_e * 1e-3), "unit": "tiles/s", "h2d_bytes_per_step": in_bytes,
                      "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": ms_e / Ke}
        del host, stages
    del store
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    _protect_stdout()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    c = make_ctx(args)
    peak, peak_src = load_peak()

    if args.workload == "slide":
        sampler = ClockSampler(c.local)
        sampler.start()
        s = run_slide(args, wl, c, args.steps, args.warmup, want_e2e=not args.no_e2e,
                      want_cpu_merge=not args.no_cpu_baseline)
        clocks = sampler.stop()
        # roofline of the dominant kernel on this workload's tiles: measured on a tiles1024 batch in the same process
        targs = argparse.Namespace(**vars(args))
        targs.workload, targs.masks, targs.steps, targs.warmup, targs.no_e2e = "tiles1024", "none", 50, 5, True
        t = run_tiles(targs, WORKLOADS["tiles1024"], c)
        line = {
            "metric": "postproc_tiles_per_s", "value": s["tiles_per_s"], "unit": "tiles/s", "n_gpus": c.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["ms_per_slide"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "slide", "slide_px": s["slide_px"], "tile": s["tile"], "overlap": s["overlap"],
                       "tiles": s["tiles"], "detections": s["detections"], "kept": s["kept"],
                       "stages": "per-tile decode+filter+compact, nms, select; append in slide coordinates; exact "
                                 "slide-level merge NMS (seam all-gather + verdict exchange over NCCL when N>1)",
                       "l2": f"{s['input_bytes'] / 1e9:.1f} GB of head outputs resident in HBM, each read once per step",
                       "launch": f"tile batches of {wl['bs']} alternate over {args.slide_streams} streams "
                                 "(SlidePostprocessor(streams=)), appends chained by events"},
            "boxes_per_s": s["boxes_per_s"], "roofline": t["roofline"], "stages": t["stages"],
            "slide": s, "e2e": s.get("e2e"), "gpu_launches": s["gpu_launches"], "clocks": clocks,
        }
    else:
        line = run_tiles(args, wl, c)
        if not args.no_slide:
            sargs = argparse.Namespace(**vars(args))
            line["slide"] = run_slide(sargs, WORKLOADS["slide"], c, steps=3, warmup=2, want_e2e=False,
                                      want_cpu_merge=not args.no_cpu_baseline)

    cpu = None
    try:   # the CPU legs use every core the process started with, not just the GPU's NUMA node
        os.sched_setaffinity(0, _ORIG_AFFINITY)
    except Exception:
        pass
    if c.rank == 0 and not args.no_cpu_baseline:
        masks = args.masks if args.workload != "slide" else "none"
        v, cores, sample = time_cpu(wl, masks, budget_s=15.0, max_reps=30)
        cpu = {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample}
        # the reference caps its own thread pool at min(8, ncpu - 1) (metayolo/__init__.py:21,31): second data point
        v8, c8, _ = time_cpu(wl, masks, budget_s=6.0, max_reps=10, threads=min(8, max(1, (os.cpu_count() or 2) - 1)))
        cpu["value_reference_threads"] = v8
        cpu["reference_threads"] = c8
    line["cpu_baseline"] = cpu
    if isinstance(line.get("e2e"), dict):
        line["e2e"]["host_affinity"] = c.affinity
    if c.rank == 0:
        _emit(line)
    if c.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
