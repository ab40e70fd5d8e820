#!/usr/bin/env bash
# round 2, first GPU pass: parity tests of everything that changed, then the slide bench small and full
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
python bench.py --slide-size 30000 --steps 2 --warmup 2 --no-cpu-baseline --no-sub --no-torch-cuda > gpurun_out/r2a_slide30k.json 2> gpurun_out/r2a_slide30k.err
tail -3 gpurun_out/r2a_slide30k.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_slide_full.json 2> gpurun_out/r2a_slide_full.err
tail -3 gpurun_out/r2a_slide_full.err
nvidia-smi --query-gpu=memory.used --format=csv
