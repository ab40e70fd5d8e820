#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_masks.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2h_pytest.txt
tail -3 gpurun_out/r2h_pytest.txt
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2h_slide.json 2> gpurun_out/r2h_slide.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda --slide-streams 2 > gpurun_out/r2h_slide_s2.json 2> gpurun_out/r2h_slide_s2.err
tail -2 gpurun_out/r2h_slide.err
