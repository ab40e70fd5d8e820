#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_next.py -x -q -m gpu 2>&1 | tail -3
for smem in 0 92160; do
  HDY_PATCH_SMEM=$smem python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2k_slide_$smem.json 2> gpurun_out/r2k_slide_$smem.err
done
HDY_PATCH_SMEM=92160 python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda --slide-streams 2 > gpurun_out/r2k_slide_92160_s2.json 2> gpurun_out/r2k_slide_92160_s2.err
