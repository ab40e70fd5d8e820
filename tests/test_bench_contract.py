"""CPU test of the bench contract that does not need a GPU: `bench.py --impl reference` (the reference's CPU path,
timed through the oracle port) prints exactly ONE JSON line on stdout with the keys the driver reads, and ranks other
than 0 print nothing."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "postproc_tiles_per_s" and d["unit"] == "tiles/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "slide" and d["config"]["masks"] == "paste"    # the reference's own mask path
    # nothing is extrapolated: ms_per_step is the measured time of one sample pass, value follows from it
    assert abs(d["value"] - d["tiles_per_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"] + 1e-9
    assert d["steps"] >= 1 and d["other_mask_variant"]["masks"] == "proto"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cpu_merge_scaling_leg():
    """The slide bench's CPU merge baseline: sub-slides are cut by tile index, every size is an exact
    Ensemble.merge of its boxes, the exponent is fitted over the sizes that fit the budget."""
    import importlib.util

    import torch
    import torchvision

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    g = torch.Generator().manual_seed(0)
    n_cols, per = 5, 300
    boxes, scores, tile = [], [], []
    for t in range(n_cols * n_cols):
        r, cc = divmod(t, n_cols)
        c = torch.rand((per, 2), generator=g) * 1024 + torch.tensor([cc * 960., r * 960.])
        s = 12 + 24 * torch.rand((per, 2), generator=g)
        boxes.append(torch.cat([c - s / 2, c + s / 2], 1))
        scores.append(0.2 + 0.7 * torch.rand(per, generator=g))
        tile.append(torch.full((per,), t))
    boxes, scores, tile = torch.cat(boxes), torch.cat(scores), torch.cat(tile)
    out = bench.cpu_merge_scaling(boxes, scores, tile, n_cols, 0.25, 0.45, grids=(2, 3, 4))
    assert out["tiles"] == [4, 9, 16] and out["boxes"] == [4 * per, 9 * per, 16 * per] and "exponent" in out
    sel = ((tile // n_cols) < 2) & ((tile % n_cols) < 2)
    b, sc = boxes[sel], scores[sel]
    keep = sc > 0.25
    assert out["kept"][0] == len(torchvision.ops.nms(b[keep], sc[keep], 0.45))
