#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/l_pytest.txt; tail -3 gpurun_out/l_pytest.txt
python bench.py --steps 5 --warmup 3 --no-sub --no-cpu-baseline --no-torch-cuda --no-e2e > gpurun_out/l_slide.json 2> gpurun_out/l_slide.err
tail -c 1500 gpurun_out/l_slide.json
