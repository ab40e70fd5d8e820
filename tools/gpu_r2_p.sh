#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_next.py -q -m gpu -k "roi" 2>&1 | tail -12 | cut -c1-250
timeout 300 python tools/roi_bench.py > gpurun_out/y_roi.json 2> gpurun_out/y_roi.err; cat gpurun_out/y_roi.json; tail -5 gpurun_out/y_roi.err
