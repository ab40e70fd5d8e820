"""End-to-end GPU test of the whole-slide pipeline (hd_yolo_b200.SlidePostprocessor): sliding_window_scanner tiles of
one synthetic nuclei field -> per-tile post-processing -> append in slide coordinates -> slide-level merge, against the
oracle composition nms_per_image + select_scores (on the device-decoded rows, so that last-ulp expf differences do not
move thresholds) -> merge_outputs -> Ensemble.merge (yolo_head.py:301-355, 450-463; yolo.py:165-204)."""
import pytest
import torch

import hd_yolo_b200 as hdy
from hd_yolo_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,tile,overlap,batch,streams", [((1500, 2000), 512, 64, 5, 1), ((900, 900), 512, 32, 128, 1),
                                                             ((1500, 2000), 512, 64, 3, 3), ((2100, 1100), 512, 64, 2, 2)])
def test_slide_postprocessor_matches_oracle_composition(cuda_device, size, tile, overlap, batch, streams):
    dev = cuda_device
    conf, iou, md = 0.25, 0.45, 1500
    spec = hdy.HeadSpec(synth.ANCHORS_3, synth.STRIDES_3, nc=4)
    post = hdy.SlidePostprocessor(spec, size, (tile, tile), overlap, conf, iou, md, cap=2048, batch=batch, device=dev,
                                  streams=streams)
    rois = post.rois
    n_tiles = len(rois)
    assert n_tiles >= 4
    store = {}

    def provider(a, b):
        if (a, b) not in store:
            store[(a, b)] = synth.slide_tile_logits(rois[a:b], tile, 4, seed=a, device=dev, pitch=22.0)
        return store[(a, b)]

    res = post.run(provider, ordered=True)
    if streams > 1:      # batches on alternating streams: a second pass over the same (now cached) inputs is identical
        again = post.run(provider, ordered=True)
        assert all(torch.equal(res[k], again[k]) for k in ('boxes', 'scores', 'labels', 'index', 'state'))
    # oracle on the same head outputs
    tiles = []
    for (a, b), dets in sorted(store.items()):
        cat = hdy.decode_concat(dets, spec).cpu()
        outs = port.nms_per_image(cat, 4, conf, iou, md)
        for j, o in enumerate(outs):
            s, l = port.select_scores(o['scores'].clone(), conf, port.default_descendants(4))
            tiles.append({'boxes': o['boxes'], 'scores': s, 'labels': l, 'roi': rois[a + j]})
    merged = port.merge_outputs(tiles)
    ref = port.ensemble_merge([{'det': merged}], {'conf_thres': conf, 'iou_thres': iou, 'max_det': 10 ** 9})['det']
    assert int(res['n']) == len(merged['boxes'])
    assert len(ref['boxes']) < len(merged['boxes'])          # the overlap bands really held duplicates
    assert torch.equal(res['boxes'].cpu(), ref['boxes'])
    assert torch.equal(res['scores'].cpu(), ref['scores'])
    assert torch.equal(res['labels'].cpu(), ref['labels'])
    # 'index' points into the merge_outputs concatenation
    assert torch.equal(merged['boxes'][res['index'].cpu()], ref['boxes'])
