#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_slide.py -q -m gpu 2>&1 | tail -2
for p in 1 0; do
HDY_NMS_PRIORITY=$p python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline --no-torch-cuda --no-e2e > gpurun_out/l_slide_$p.json 2> gpurun_out/l_slide.err
python - <<PY
import json
d=json.loads(open('gpurun_out/l_slide_$p.json').read().strip().splitlines()[-1])
print('prio $p', d['ms_per_step'], {k:round(d['slide'][k],2) for k in ('detect_ms','merge_ms','masks_ms')}, d['slide']['digest']['hash'], d['slide']['mask_digest']['hash'])
PY
done
