#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_next.py -q -m gpu -k "roi" 2>&1 | tail -3
timeout 300 python tools/roi_bench.py > gpurun_out/x_roi.json 2> gpurun_out/x_roi.err; cat gpurun_out/x_roi.json; tail -5 gpurun_out/x_roi.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_tc_kernel -s 3 -c 1 -o gpurun_out/x_roi_tc python tools/roi_bench.py 8192 > gpurun_out/x_ncu.log 2>&1
tail -2 gpurun_out/x_ncu.log
