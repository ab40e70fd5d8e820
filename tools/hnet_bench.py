"""bench.py --workload hnet (BASELINE.json configs[4]): hnet multi-level heads over one synthetic slide at two
magnifications, both sharded over the ranks by the same spatial row bands:

  40x  nuclei tiles (1024 px, 64 px overlap)      YOLO decode + class-agnostic per-tile NMS + score select, slide merge
                                                   (SlidePostprocessor: the configs[3] path without masks)
  10x  structure tiles (1024 px at 1/4 resolution) RCNN-style heads exactly as hnet/detection/mask_rcnn.py:41-75, 145-298
                                                   delegates them to torchvision: BoxCoder.decode (H1), RPN
                                                   filter_proposals with per-level top-k + batched_nms 0.7 (H2),
                                                   RoIHeads.postprocess_detections: softmax, per-class decode, score
                                                   cut, class-aware batched_nms 0.5, top-100 (H3; thresholds
                                                   hnet/detection/utils_det.py:16-52)
  cross-level  rescale_outputs(scale=4) -> merge_outputs -> per-task Ensemble.merge (yolo_head.py:450-471,
               yolo.py:165-204; the hnet-side composition is a TODO in the reference, hnet/hnet_new.py:275): the
               structure detections of all ranks are merged through the same seam exchange as the nuclei.

One "step" = the whole two-magnification slide.  Imported by bench.py (run_hnet)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_hnet(args, wl, c, common_config, load_peak, ClockSampler):
    import torch
    import hd_yolo_b200 as hdy
    from hd_yolo_b200 import dist as hdist
    from hd_yolo_b200 import hnet, ops, synth, synth_hnet
    from hd_yolo_b200.pipeline import SlidePostprocessor
    from hd_yolo_b200.slide import (ensemble_merge, fold_digest, kept_digest, merge_outputs, rescale_outputs,
                                    sliding_window_scanner)

    dev, S, tile, nc = c.dev, args.slide_size, wl["tile"], wl["nc"]
    conf, iou = wl["conf"], wl["iou"]
    # ---- 40x: the slide path without masks -------------------------------------------------------------------
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=nc)
    bs40 = 148
    post = SlidePostprocessor(spec, (S, S), (tile, tile), wl["overlap"], conf, iou, 3328, cap=wl["cap"], batch=bs40,
                              rank=c.rank, world=c.world, device=dev, streams=args.slide_streams)
    t0, t1 = post.tile_range
    store = [synth.slide_tile_logits(post.rois[a:min(a + bs40, t1)], tile, nc, seed=1, first_tile=a, conf=conf,
                                     device=dev) for a in range(t0, t1, bs40)]
    in40 = sum(sum(t.numel() * 4 for t in d) for d in store)

    # ---- 10x: RCNN heads ---------------------------------------------------------------------------------------
    S10 = S // 4
    rois10 = sliding_window_scanner((S10, S10), (tile, tile), wl["overlap"])
    u0, u1 = hdist.shard_tile_rows(rois10, c.world)[c.rank]
    bs10, C10 = wl["bs"], 3                                    # background + nodule + secondary structure
    anchors, counts = synth_hnet.pyramid_anchors(tile)
    anchors = anchors.to(dev)
    A = int(anchors.shape[0])
    shapes = [(tile, tile)] * bs10
    rpn_in = []
    for a in range(u0, u1, bs10):
        n = min(a + bs10, u1) - a
        g = torch.Generator(device=dev).manual_seed(50000 + a)
        rpn_in.append((torch.randn((n, A), generator=g, device=dev) * 2.0,                 # objectness logits
                       torch.randn((n * A, 4), generator=g, device=dev) * 0.1,             # deltas N(0, 0.1)
                       torch.randn((n * 1000, C10), generator=g, device=dev),              # class logits N(0, 1)
                       torch.randn((n * 1000, C10 * 4), generator=g, device=dev) * 0.1))   # box regression
    in10 = sum(sum(t.numel() * 4 for t in b) for b in rpn_in)
    comm = hdist.TorchDistComm() if c.world > 1 else None
    params = {'conf_thres': 0.05, 'iou_thres': 0.5, 'max_det': 10 ** 9}
    res = {}

    def step(i):
        # 40x nuclei (task "nuclei")
        res["nuclei"] = post.run(lambda a, b: store[(a - t0) // bs40], ordered=True)
        # 10x structures (task "structure"): per-tile heads, then the cross-level merge in the 40x frame
        tiles = []
        for k, (obj, deltas, cl, br) in enumerate(rpn_in):
            a = u0 + k * bs10
            n = obj.shape[0]
            prop = hnet.box_decode(deltas, anchors, boxes_rows=A).view(n, A, 4)                       # H1
            pb, _ = hnet.rpn_filter_proposals(prop, obj, shapes[:n], counts, 1000, 1000, 0.7, 0.0,
                                               mode="vanilla")                                        # H2
            R = sum(int(p.shape[0]) for p in pb)
            db, ds, dl = hnet.roi_postprocess_detections(cl[:R], br[:R], pb, shapes[:n], score_thresh=0.05,
                                                         nms_thresh=0.5, detections_per_img=100, mode="vanilla")  # H3
            for j in range(n):
                tiles.append({'boxes': db[j], 'scores': ds[j], 'labels': dl[j], 'roi': rois10[a + j]})
        if tiles:
            m = rescale_outputs(merge_outputs(tiles), 4.0)          # 10x -> 40x frame (yolo_head.py:465-471)
        else:
            m = {'boxes': torch.zeros((0, 4), device=dev), 'scores': torch.zeros((0,), device=dev),
                 'labels': torch.zeros((0,), dtype=torch.int64, device=dev)}
        if c.world > 1:
            r = hdist.merge_sharded(m['boxes'].contiguous(), m['scores'].contiguous(), params['conf_thres'],
                                    params['iou_thres'], comm=comm, seam_cap=8192)
            res["structure"] = {'state': r['state'], 'base': r['base'], 'n': int(m['scores'].shape[0])}
        else:
            out = ensemble_merge([{'structure': m}], params)['structure']
            st = torch.zeros((m['scores'].shape[0],), dtype=torch.uint8, device=dev)
            res["structure"] = {'kept': int(out['scores'].shape[0]), 'n': int(m['scores'].shape[0])}

    sampler = ClockSampler(c.local)
    sampler.start()
    for i in range(max(args.warmup, 2)):
        step(i)
    ops.profile.reset()
    ms = c.timed(step, args.steps)
    launches = ops.profile.launches
    clocks = sampler.stop()
    ops.profile.enabled = True
    ops.profile.reset()
    step(0)
    prof = ops.profile.summary()
    ops.profile.enabled = False
    n10 = u1 - u0
    alg = {   # this rank's algorithmic bytes per step for the hnet calls
        "hdy_rcnn_decode": n10 * A * (16 + 16) + A * 16,
        "hdy_rpn_level_keys": n10 * A * (4 + 8),
        "hdy_sort_keys_bytes": n10 * A * 8 * 2 * 8,
        "hdy_softmax_rows": n10 * 1000 * C10 * 8,
        "hdy_filter_compact_logits": in40,
    }
    stages = {}
    for k, (n, t) in prof.items():
        a_b = alg.get(k)
        stages[k] = {"calls": n, "ms_total": t, "ms": t / max(n, 1), "alg_bytes_total": a_b,
                     "gbs": (a_b / (t * 1e-3) / 1e9) if a_b and t > 0 else None}
    nuc = res["nuclei"]
    tot = torch.cat([torch.tensor([int(nuc["n"]), res["structure"]["n"], in40 + in10], dtype=torch.int64, device=dev),
                     kept_digest(nuc["state"], nuc["base"])])
    if c.world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot)
    n_tiles = int(post.rois.shape[0]) + int(rois10.shape[0])
    peak, peak_src = load_peak()
    dom = max((k for k in stages if stages[k]["gbs"]), key=lambda k: stages[k]["ms_total"])
    per_launch = alg[dom] / stages[dom]["calls"]
    ach = stages[dom]["gbs"]
    cfg = common_config(args, wl, "none")
    cfg.update({"slide_px": S, "tiles_40x": int(post.rois.shape[0]), "tiles_10x": int(rois10.shape[0]),
                "anchors_per_10x_tile": A, "classes_10x": C10, "nuclei_detections": int(tot[0]),
                "structure_detections": int(tot[1]), "nuclei_digest": fold_digest(tot[3:6]),
                "stages": "40x: decode+filter+compact, nms, select, slide merge; 10x: BoxCoder.decode, RPN "
                          "filter_proposals (per-level top-k, batched_nms 0.7), RoI postprocess_detections (softmax, "
                          "class-aware batched_nms 0.5, top-100); cross-level: rescale x4, merge_outputs, per-task "
                          "Ensemble.merge (seam exchange over NCCL when N>1)",
                "l2": f"{int(tot[2]) / 1e9:.1f} GB of head outputs resident in HBM, each read once per step"})
    return {
        "metric": "postproc_tiles_per_s", "value": n_tiles * args.steps / (ms * 1e-3), "unit": "tiles/s",
        "n_gpus": c.world, "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "stages": stages,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch,
                     "kernel_ms": stages[dom]["ms"]},
        "e2e": None, "gpu_launches": launches, "clocks": clocks,
    }
