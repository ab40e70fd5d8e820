"""CPU tests of the sharded whole-slide merge's host logic (hd_yolo_b200/dist.py): tile sharding, the fixed-size
seam blocks and their four all-gathers over gloo (world_size 2 and 3) and over thread-emulated ranks, verdict
exchange, termination, and block growth after an overflow.  The kernels are replaced by the dense stand-in in
tests/cpu_merge_backend.py; the checker is torchvision.ops.nms on the slide-wide concatenation, which is what the
reference's Ensemble.merge calls (metayolo/models/yolo.py:189-195)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp
import torchvision

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cpu_merge_backend import TorchSeamBackend  # noqa: E402
from slide_synth import banded_detections  # noqa: E402

from hd_yolo_b200 import dist as hdist  # noqa: E402
from hd_yolo_b200.slide import sliding_window_scanner  # noqa: E402

CONF, IOU = 0.25, 0.45


def reference_states(parts, conf=CONF, iou=IOU):
    """Ensemble.merge verdicts on the concatenation: 1 kept, 2 suppressed, 3 dropped."""
    boxes = torch.cat([p[0] for p in parts])
    scores = torch.cat([p[1] for p in parts])
    keep = scores > conf
    idx = torch.nonzero(keep).flatten()
    kept = idx[torchvision.ops.nms(boxes[idx], scores[idx], iou)]
    st = torch.full((len(scores),), 2, dtype=torch.uint8)
    st[~keep] = 3
    st[kept] = 1
    return st


def test_shard_tile_rows_covers_every_tile_once():
    rois = sliding_window_scanner((10000, 7000), (1024, 1024), 64)
    n_cols = len(torch.unique(rois[:, 0]))
    for world in (1, 2, 3, 8, 16):
        sh = hdist.shard_tile_rows(rois, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(rois)
        for (a, b), (c, d) in zip(sh[:-1], sh[1:]):
            assert b == c
        for a, b in sh:
            assert (b - a) % n_cols == 0          # whole tile rows
        sizes = [(b - a) // n_cols for a, b in sh]
        assert max(sizes) - min(sizes) <= 1


def test_shard_tile_rows_full_slide_config():
    rois = sliding_window_scanner((100000, 100000), (1024, 1024), 64)
    assert len(rois) == 11025
    sh = hdist.shard_tile_rows(rois, 8)
    assert [(b - a) // 105 for a, b in sh] == [13, 13, 13, 13, 13, 13, 13, 14]


@pytest.mark.parametrize("world", [2, 3, 5])
def test_emulated_ranks_match_dense_nms(world):
    parts = banded_detections(world, seed=world, n_nuclei=300 * world)
    ref = reference_states(parts)
    got = hdist.merge_emulated(parts, CONF, IOU, backend=TorchSeamBackend)
    assert torch.equal(torch.cat(got), ref)
    # the exchange really carried something: some detection was suppressed by another rank's box
    alone = torch.cat([reference_states([p]) for p in parts])
    assert (alone != ref).any()


def test_emulated_empty_and_single_rank():
    parts = banded_detections(3, seed=9, n_nuclei=200)
    parts[1] = (torch.zeros((0, 4)), torch.zeros((0,)))
    got = hdist.merge_emulated(parts, CONF, IOU, backend=TorchSeamBackend)
    assert torch.equal(torch.cat(got), reference_states(parts))
    one = hdist.merge_emulated(parts[:1], CONF, IOU, backend=TorchSeamBackend)
    assert torch.equal(one[0], reference_states(parts[:1]))


def test_payload_overflow_grows_the_blocks():
    """seam_cap far too small: every rank sees the overflow flag in the same read, grows and repeats."""
    parts = banded_detections(3, seed=4, n_nuclei=900)
    ref = reference_states(parts)
    got = hdist.merge_emulated(parts, CONF, IOU, backend=TorchSeamBackend, seam_cap=8)
    assert torch.equal(torch.cat(got), ref)


def test_emulated_failure_does_not_deadlock():
    def boom(rank, comm):
        if rank == 1:
            raise ValueError("rank 1 fails")
        comm.all_gather(torch.zeros(1))
    with pytest.raises(ValueError):
        hdist.run_emulated(3, boom)


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = banded_detections(world, seed=11, n_nuclei=250 * world)
        b, s = parts[rank]
        res = hdist.merge_sharded(b, s, CONF, IOU, backend=TorchSeamBackend, seam_cap=64 if world == 3 else 4096)
        out[rank] = (res['state'].clone(), res['base'], res['exchanges'], res['seam_rows'])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_match_dense_nms(world):
    port = 29500 + os.getpid() % 2000 + world
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, out), nprocs=world, join=True)
        parts = banded_detections(world, seed=11, n_nuclei=250 * world)
        ref = reference_states(parts)
        got = torch.cat([out[r][0] for r in range(world)])
        assert torch.equal(got, ref)
        bases = [out[r][1] for r in range(world)]
        assert bases == [sum(len(p[1]) for p in parts[:r]) for r in range(world)]
        assert all(out[r][2] >= 1 for r in range(world))
        assert sum(out[0][3]) > 0
