#!/usr/bin/env python
"""Benchmark of the post-processing hot path (BASELINE.json metric: post-proc tiles/s & boxes/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tiles640|tiles1024]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path (fused decode+filter+compact -> per-tile NMS -> score/label
select) over one batch of synthetic head outputs.  Prints ONE JSON line on rank 0.

  value     : whole-job tiles/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       : the same through the public API with HOST (pinned) inputs: H2D of the step's logits and
              D2H of its detections inside the timed region
  roofline  : dominant kernel (hdy_filter_compact_logits) algorithmic bytes / its CUDA-event time
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

WORKLOADS = {
    # BASELINE.json configs[1]: 640x640 tiles, batch 64, 3 anchor levels, ~1k candidates/tile
    "tiles640": dict(tile=640, bs=64, n_cand=1000, conf=0.25, iou=0.45, max_det=1000, nc=4, cap=2048),
    # BASELINE.json configs[2]: 1024x1024 dense-nuclei tiles, batch 128, ~3k candidates/tile
    "tiles1024": dict(tile=1024, bs=128, n_cand=3000, conf=0.25, iou=0.45, max_det=3000, nc=4, cap=4096),
}
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiles640", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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


def cpu_reference_pass(dets_cpu, wl, spec_args):
    """The reference's CPU path for one batch: compute_proposals -> pad/cat -> nms_per_image -> select."""
    from oracle import port
    anchors, strides = spec_args
    preds = port.compute_proposals(dets_cpu, anchors, strides)
    params = {'conf_thres': wl["conf"], 'iou_thres': wl["iou"], 'max_det': wl["max_det"]}
    return port.compute_outputs(preds, wl["nc"], params)


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation (oracle port: same torch/torchvision
    calls) on all host threads; each step is a bounded sample of the workload."""
    import torch
    from hd_yolo_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(wl["bs"], 8)
    torch.set_num_threads(os.cpu_count() or 1)
    dets = synth.nuclei_logits(sample, wl["tile"], wl["nc"], wl["n_cand"], seed=1, conf=wl["conf"])
    spec_args = (synth.ANCHORS_3, synth.STRIDES_3)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_reference_pass(dets, wl, spec_args)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(dets, wl, spec_args)
    dt = time.perf_counter() - t0
    v = sample * steps / dt
    line = {
        "impl": "reference", "metric": "postproc_tiles_per_s", "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "tile": wl["tile"], "sample_tiles_per_step": sample,
                   "candidates_per_tile": wl["n_cand"], "conf": wl["conf"], "iou": wl["iou"], "max_det": wl["max_det"]},
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} tiles/step x {steps} steps, oracle/port.py (torch {torch.__version__} CPU + torchvision nms)"},
        "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "boxes_per_s": v * wl["n_cand"],
    }
    print(json.dumps(line))


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    import hd_yolo_b200 as hdy
    from hd_yolo_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"

    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=wl["nc"])
    bs, K, W = wl["bs"], args.steps, args.warmup
    shapes = synth.level_shapes(wl["tile"], synth.STRIDES_3)
    N = spec.rows_per_tile(shapes)
    in_bytes = bs * N * spec.no * 4
    R = max(2, int(2.5 * L2_BYTES / in_bytes) + 1)       # rotate input batches so that reads miss L2
    batches = [synth.nuclei_logits(bs, wl["tile"], wl["nc"], wl["n_cand"], seed=1000 * rank + r, conf=wl["conf"],
                                   generator_device="cuda") for r in range(R)]

    def step(i, dets=None):
        return hdy.detect_postprocess(dets if dets is not None else batches[i % R], spec, wl["conf"], wl["iou"],
                                      wl["max_det"], cap=wl["cap"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    sampler = ClockSampler(local)
    sampler.start()

    # ---- value: inputs resident in HBM ---------------------------------------------------------------
    for i in range(max(W, 3)):
        out = step(i)
    out.to_list()                      # raises on capacity overflow
    ops.profile.reset()
    ms = timed(step, K)
    launches = ops.profile.launches
    tiles_per_s = world * bs * K / (ms * 1e-3)

    # ---- per-kernel CUDA-event times over an identical region (roofline) --------------------------------
    ops.profile.enabled = True
    ops.profile.reset()
    timed(step, K)
    prof = ops.profile.summary()
    ops.profile.enabled = False
    cand_mean = float(out.cand_counts[:bs].float().mean())
    kept_mean = float(out.counts.float().mean())
    dom = "hdy_filter_compact_logits"
    dom_calls, dom_ms = prof[dom]
    alg_bytes = bs * (4 * N * spec.no + 24 * cand_mean)          # read logits once, write key(8)+box(16) per candidate
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (dom_ms / dom_calls * 1e-3) / 1e9
    stage_ms = {k: v[1] / v[0] for k, v in prof.items()}
    # whole-step algorithmic bytes: decode + nms read/write
    nc = wl["nc"]
    step_bytes = alg_bytes + bs * (cand_mean * 24 + kept_mean * (4 * (1 + nc)) + kept_mean * (16 + 4 + 8 + 4 + 4 * (1 + nc) + 4))

    # ---- e2e: host (pinned) inputs, H2D + D2H inside the timed region ---------------------------------------
    e2e = None
    if not args.no_e2e:
        host = [[d.cpu().pin_memory() for d in b] for b in batches[:2]]
        stage = [torch.empty_like(d) for d in batches[0]]
        md = min(wl["max_det"], wl["cap"])
        h_boxes = torch.empty((bs, md, 4), dtype=torch.float32).pin_memory()
        h_scores = torch.empty((bs, md), dtype=torch.float32).pin_memory()
        h_labels = torch.empty((bs, md), dtype=torch.int64).pin_memory()
        h_counts = torch.empty((bs,), dtype=torch.int32).pin_memory()

        def e2e_step(i):
            for s, h in zip(stage, host[i % 2]):
                s.copy_(h, non_blocking=True)
            o = step(i, stage)
            h_boxes.copy_(o.boxes, non_blocking=True)
            h_scores.copy_(o.scores, non_blocking=True)
            h_labels.copy_(o.labels, non_blocking=True)
            h_counts.copy_(o.counts, non_blocking=True)

        Ke = max(3, min(K, 100))
        for i in range(3):
            e2e_step(i)
        ms_e = timed(e2e_step, Ke)
        e2e = {"value": world * bs * Ke / (ms_e * 1e-3), "unit": "tiles/s", "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": h_boxes.numel() * 4 + h_scores.numel() * 4 + h_labels.numel() * 8 + h_counts.numel() * 4,
               "steps": Ke, "ms_per_step": ms_e / Ke}

    clocks = sampler.stop()

    # ---- CPU baseline: oracle port on host cores, rank 0, bounded sample -----------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sample = min(bs, 16)
        dets_cpu = [d[:sample].cpu() for d in batches[0]]
        spec_args = (synth.ANCHORS_3, synth.STRIDES_3)
        cpu_reference_pass(dets_cpu, wl, spec_args)
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 50):
            cpu_reference_pass(dets_cpu, wl, spec_args)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": sample * reps / dt, "unit": "tiles/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} tiles x {reps} passes of oracle/port.py (torch CPU + torchvision.ops.nms)"}

    if rank == 0:
        line = {
            "metric": "postproc_tiles_per_s", "value": tiles_per_s, "unit": "tiles/s", "n_gpus": world, "steps": K,
            "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "tile": wl["tile"], "tiles_per_step_per_gpu": bs, "levels": 3,
                       "rows_per_tile": N, "channels": spec.no, "candidates_per_tile": round(cand_mean, 1),
                       "kept_per_tile": round(kept_mean, 1), "conf": wl["conf"], "iou": wl["iou"],
                       "max_det": wl["max_det"], "cap": wl["cap"], "stages": "decode+filter+compact, nms, score/label select",
                       "l2": f"{R} rotating input batches ({R * in_bytes / 1e6:.0f} MB) > 126 MB L2"},
            "boxes_per_s": tiles_per_s * cand_mean,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms / dom_calls},
            "pipeline": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms / K * 1e-3) / 1e9 * 1.0,
                         "frac_of_peak": step_bytes / (ms / K * 1e-3) / 1e9 / peak, "stage_ms": stage_ms},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
