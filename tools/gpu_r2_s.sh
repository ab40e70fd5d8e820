#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_masks.py tests/test_gpu_pipeline.py tests/test_gpu_core.py -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline --no-torch-cuda --no-e2e > gpurun_out/l_slide.json 2> gpurun_out/l_slide.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/l_slide.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], {k:d['slide'][k] for k in ('detect_ms','merge_ms','masks_ms')}, d['slide']['digest'], d['slide']['mask_digest'])
PY
