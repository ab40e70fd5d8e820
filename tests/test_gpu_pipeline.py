"""End-to-end GPU tests of the whole-slide pipeline (hd_yolo_b200.SlidePostprocessor): sliding_window_scanner tiles of
one synthetic nuclei field -> per-tile post-processing -> append in slide coordinates -> slide-level merge (-> masks of
the kept rows), against the oracle composition nms_per_image + select_scores (on the device-decoded rows, so that
last-ulp expf differences do not move thresholds) -> merge_outputs -> Ensemble.merge (yolo_head.py:301-355, 450-463;
yolo.py:165-204), and -- for W ranks emulated through the REAL multi-rank code path of pipeline.py / dist.py, interior
shortcut on -- against the single-rank run, row by row."""
import pytest
import torch

import hd_yolo_b200 as hdy
from hd_yolo_b200 import dist as hdist
from hd_yolo_b200 import synth
from hd_yolo_b200.ops import scratch_slot
from hd_yolo_b200.slide import fold_digest, kept_digest, mask_digest
from oracle import port

pytestmark = pytest.mark.gpu
CONF, IOU = 0.25, 0.45


def _oracle_merge(spec, rois, store, md):
    tiles = []
    for (a, b), dets in sorted(store.items()):
        cat = hdy.decode_concat(dets, spec).cpu()
        ne = spec.no - 5 - spec.nc
        if ne:      # mask coefficients travel RAW (the decode's sigmoid over the whole row is not applied to them)
            raw = torch.cat([d.reshape(d.shape[0], -1, spec.no) for d in dets], 1).cpu()
            cat[..., 5 + spec.nc:5 + spec.nc + ne] = raw[..., 5 + spec.nc:]
        outs = port.nms_per_image(cat, spec.nc, CONF, IOU, md)
        for j, o in enumerate(outs):
            s, l = port.select_scores(o['scores'][:, :1 + spec.nc].clone(), CONF, port.default_descendants(spec.nc))
            tiles.append({'boxes': o['boxes'], 'scores': s, 'labels': l, 'roi': rois[a + j],
                          'extra': o['extra']})
    merged = port.merge_outputs(tiles)
    ref = port.ensemble_merge([{'det': merged}], {'conf_thres': CONF, 'iou_thres': IOU, 'max_det': 10 ** 9})['det']
    return tiles, merged, ref


@pytest.mark.parametrize("size,tile,overlap,batch,streams", [((1500, 2000), 512, 64, 5, 1), ((900, 900), 512, 32, 128, 1),
                                                             ((1500, 2000), 512, 64, 3, 3), ((2100, 1100), 512, 64, 2, 2)])
def test_slide_postprocessor_matches_oracle_composition(cuda_device, size, tile, overlap, batch, streams):
    dev = cuda_device
    md = 1500
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    post = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=batch, device=dev,
                                  streams=streams)
    rois = post.rois
    n_tiles = len(rois)
    assert n_tiles >= 4
    store = {}

    def provider(a, b):
        if (a, b) not in store:
            store[(a, b)] = synth.slide_tile_logits(rois[a:b], tile, 4, seed=1, first_tile=a, device=dev, pitch=22.0)
        return store[(a, b)]

    res = post.run(provider, ordered=True)
    if streams > 1:      # batches on alternating streams: a second pass over the same (now cached) inputs is identical
        again = post.run(provider, ordered=True)
        assert all(torch.equal(res[k], again[k]) for k in ('boxes', 'scores', 'labels', 'index', 'state'))
    # oracle on the same head outputs
    _, merged, ref = _oracle_merge(spec, rois, store, md)
    assert int(res['n']) == len(merged['boxes'])
    assert len(ref['boxes']) < len(merged['boxes'])          # the overlap bands really held duplicates
    assert torch.equal(res['boxes'].cpu(), ref['boxes'])
    assert torch.equal(res['scores'].cpu(), ref['scores'])
    assert torch.equal(res['labels'].cpu(), ref['labels'])
    # 'index' points into the merge_outputs concatenation
    assert torch.equal(merged['boxes'][res['index'].cpu()], ref['boxes'])


def _plant_far_boxes(dets, where):
    """Huge, confident false positives (a few hundred px) in the stride-32 level: boxes that stick far out of their
    tile and reach into other tiles -- and, near a band boundary, into another rank's tiles (far list, dirty tiles)."""
    big = dets[2]
    for (b, gy, gx, sw) in where:
        sig = torch.tensor([0.5, 0.5, sw, sw, 0.97], device=big.device)
        big[b, 2, gy, gx, :5] = torch.log(sig / (1 - sig))


@pytest.mark.parametrize("world,far", [(2, False), (3, True), (4, True), (8, True)])   # 8 ranks, 6 tile rows: two idle
def test_sharded_ranks_match_single_rank_with_shortcut(cuda_device, world, far):
    """SlidePostprocessor(world=W, rank=r) for every r -- through pipeline.merge's multi-rank branch with the interior
    shortcut, seam blocks, far lists and dirty tiles -- gives, row for row, the verdicts of the world=1 run and of the
    oracle composition; the digests the bench prints agree too."""
    dev = cuda_device
    size, tile, overlap, md = (2600, 1700), 512, 64, 1500
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    one = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=3, device=dev)
    rois = one.rois
    store = {}

    def tile_logits(t):
        if t not in store:
            d = synth.slide_tile_logits(rois[t:t + 1], tile, 4, seed=5, first_tile=t, device=dev, pitch=21.0)
            if far and t % 5 == 2:
                _plant_far_boxes(d, [(0, 1 + t % 7, 14 - t % 5, 0.62), (0, 12, 2 + t % 9, 0.55)])
            store[t] = d
        return store[t]

    def provider(a, b):
        ts = [tile_logits(t) for t in range(a, b)]
        return [torch.cat([x[l] for x in ts]).contiguous() for l in range(3)]

    assert one.shortcut
    ref1 = one.run(provider, ordered=True)
    n_all = int(ref1['n'])

    def rank_run(rank, comm):
        with scratch_slot(100 + rank):
            post = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=2,
                                          device=dev, rank=rank, world=world, comm=comm, seam_cap=256,
                                          streams=1 + rank % 2)
            assert post.shortcut
            r = post.run(provider, ordered=True)
            r['digest'] = kept_digest(r['state'], r['base'])
            return r

    parts = hdist.run_emulated(world, rank_run)
    # rank order == tile order: the concatenation of the ranks' rows is the single-rank row order
    assert [p['base'] for p in parts] == [sum(int(q['n']) for q in parts[:i]) for i in range(world)]
    state = torch.cat([p['state'] for p in parts])
    assert state.numel() == n_all
    assert torch.equal(state, ref1['state'])
    assert fold_digest(sum(p['digest'] for p in parts)) == fold_digest(kept_digest(ref1['state'], 0))
    assert sum(sum(p['seam_rows']) for p in parts) > 0 and all(p['exchanges'] >= 2 for p in parts)
    assert world < 8 or sum(int(p['n']) == 0 for p in parts) == 2
    # survivors: every rank's list is score-descending; their union is the single-rank list
    for p in parts:
        assert bool((p['scores'][1:] <= p['scores'][:-1]).all())
    idx = torch.cat([p['index'] for p in parts])
    order = torch.argsort(idx)
    o1 = torch.argsort(ref1['index'])
    assert torch.equal(idx[order], ref1['index'][o1])
    assert torch.equal(torch.cat([p['boxes'] for p in parts])[order], ref1['boxes'][o1])
    assert torch.equal(torch.cat([p['labels'] for p in parts])[order], ref1['labels'][o1])
    # ... and the oracle composition on the same head outputs agrees with both
    per_batch = {(t, t + 1): store[t] for t in sorted(store)}
    _, merged, ref = _oracle_merge(spec, rois, per_batch, md)
    assert torch.equal(ref1['boxes'].cpu(), ref['boxes']) and torch.equal(ref1['scores'].cpu(), ref['scores'])
    if far:
        assert float((merged['boxes'][:, 2] - merged['boxes'][:, 0]).max()) > 200.0


def test_slide_masks_of_kept_rows_match_oracle(cuda_device):
    """Masks in the slide: process_mask runs after the slide-level verdicts, on KEPT rows only, bit-packed in slide
    pixels.  Checked against the oracle's process_mask (upsample, > 0.5) per tile on the same boxes / coefficients /
    prototypes: >= 99.99 % of the pixels, and rows that were not kept own no words.  Sharded (W = 2) masks equal the
    single-rank ones bit for bit (same digest)."""
    dev = cuda_device
    size, tile, overlap, md, nm = (700, 900), 512, 64, 1200, 32
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4, no=9 + nm)
    post = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=4, device=dev)
    rois = post.rois
    store, pstore = {}, {}

    def provider(a, b):
        if (a, b) not in store:
            store[(a, b)] = synth.slide_tile_logits(rois[a:b], tile, 4, seed=9, first_tile=a, device=dev, pitch=30.0,
                                                    extra=nm)
        return store[(a, b)]

    def protos(a, b):
        if (a, b) not in pstore:
            pstore[(a, b)] = synth.slide_tile_protos(b - a, tile, seed=9, first_tile=a, nm=nm, device=dev)
        return pstore[(a, b)]

    res = post.run(provider, ordered=True, proto_provider=protos)
    pm = res['masks']
    pm.check()
    state = res['state'].cpu()
    n = int(res['n'])
    H, W = pm.H, pm.W                                             # canvas: the image padded to whole tiles
    dense = pm.to_dense().cpu()                                    # [n, H, W] uint8 in slide pixels
    words = ((pm.geom[:, 2] + 31) // 32 * pm.geom[:, 3]).cpu()
    assert int(words[state != 1].sum()) == 0                      # suppressed / dropped rows own no words
    assert int(words[state == 1].sum()) == int(pm.offsets[n])
    tiles, merged, ref = _oracle_merge(spec, rois, store, md)
    row = 0
    agree, total = 0, 0
    all_protos = torch.cat([pstore[k] for k in sorted(pstore)]).cpu()
    for t, td in enumerate(tiles):
        k = len(td['boxes'])
        if k == 0:
            continue
        refm = port.process_mask(all_protos[t], td['extra'][:, :nm], td['boxes'].clone(), (tile, tile), upsample=True)
        x0, y0 = int(rois[t, 0]), int(rois[t, 1])
        hh, ww = min(tile, H - y0), min(tile, W - x0)
        for j in range(k):
            if state[row + j] == 1:
                got = dense[row + j, y0:y0 + hh, x0:x0 + ww].float()
                want = refm[j, :hh, :ww]
                agree += int((got == want).sum())
                total += hh * ww
                # nothing outside the tile's own window
                assert int(dense[row + j].sum()) == int(got.sum())
        row += k
    assert row == n and total > 0
    assert agree / total >= 0.9999, f"mask agreement {agree / total}"
    set_px = int(dense[state == 1].sum())
    assert set_px > 1e5 and total - agree <= 1e-4 * set_px, f"{total - agree} of {set_px} set pixels differ"
    d1 = fold_digest(mask_digest(pm, res['state'], 0))

    def rank_run(rank, comm):
        with scratch_slot(200 + rank):
            p2 = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=3,
                                        device=dev, rank=rank, world=2, comm=comm, seam_cap=512)

            def prov(a, b):
                return [torch.cat([provider(t, t + 1)[l] for t in range(a, b)]) for l in range(3)]

            def prot(a, b):
                return torch.cat([protos(t, t + 1) for t in range(a, b)])
            r = p2.run(prov, ordered=False, proto_provider=prot)
            r['masks'].check()
            return mask_digest(r['masks'], r['state'], r['base']), kept_digest(r['state'], r['base'])

    # per-tile stores for the sharded run (batch boundaries differ between the runs; the synth is per tile)
    store.clear()
    pstore.clear()
    parts = hdist.run_emulated(2, rank_run)
    assert fold_digest(parts[0][0] + parts[1][0]) == d1
    assert fold_digest(parts[0][1] + parts[1][1]) == fold_digest(kept_digest(res['state'], 0))


@pytest.mark.parametrize("world", [1, 2])
def test_run_with_side_streams_equals_one_stream(cuda_device, world):
    """run(ordered=True, proto_provider=...) with three side streams (tile batches and mask batches alternate over
    them) gives the survivors in the same order and the same mask bits as the one-stream sequence, twice in a row."""
    dev = cuda_device
    size, tile, overlap, md, nm = (1500, 1100), 512, 64, 1200, 32
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4, no=9 + nm)
    rois = hdy.sliding_window_scanner(size, (tile, tile), overlap)
    lg = [synth.slide_tile_logits(rois[t:t + 1], tile, 4, seed=13, first_tile=t, device=dev, pitch=30.0, extra=nm)
          for t in range(len(rois))]
    pr = [synth.slide_tile_protos(1, tile, seed=13, first_tile=t, nm=nm, device=dev) for t in range(len(rois))]

    def prov(a, b):
        return [torch.cat([lg[t][l] for t in range(a, b)]) for l in range(3)]

    def prot(a, b):
        return torch.cat(pr[a:b])

    def rank_run(rank, comm, streams):
        with scratch_slot(300 + 10 * streams + rank):
            p = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, CONF, IOU, md, cap=2048, batch=2, device=dev,
                                       rank=rank, world=world, comm=comm, seam_cap=512, streams=streams)
            outs = []
            for _ in range(2):          # twice: the second run re-uses the streams and scratch of the first
                r = p.run(prov, ordered=True, proto_provider=prot)
                r['masks'].check()
                outs.append((r['index'].clone(), r['boxes'].clone(), r['scores'].clone(), r['labels'].clone(),
                             mask_digest(r['masks'], r['state'], r['base']), r['state'].clone()))
            for a, b in zip(outs[0], outs[1]):
                assert torch.equal(a, b)
            return outs[1]

    if world == 1:
        one, three = rank_run(0, None, 1), rank_run(0, None, 3)
        parts1, parts3 = [one], [three]
    else:
        parts1 = hdist.run_emulated(world, lambda r, c: rank_run(r, c, 1))
        parts3 = hdist.run_emulated(world, lambda r, c: rank_run(r, c, 3))
    for a, b in zip(parts1, parts3):
        assert a[0].numel() > 100
        for x, y in zip(a, b):
            assert torch.equal(x, y)
