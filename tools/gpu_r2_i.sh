#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2i_pytest.txt
tail -3 gpurun_out/r2i_pytest.txt
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2i_slide.json 2> gpurun_out/r2i_slide.err
HDY_NMS_FULL=1 python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2i_slide_full.json 2> gpurun_out/r2i_slide_full.err
tail -2 gpurun_out/r2i_slide.err
python bench.py --workload hnet --slide-size 40000 --steps 2 --warmup 2 --no-cpu-baseline --no-torch-cuda > gpurun_out/r2i_hnet.json 2> gpurun_out/r2i_hnet.err
tail -5 gpurun_out/r2i_hnet.err
