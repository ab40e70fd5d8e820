#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_next.py -q -m gpu -k "roi" 2>&1 | tail -40 > gpurun_out/m_pytest.txt; tail -30 gpurun_out/m_pytest.txt
timeout 300 python tools/roi_bench.py > gpurun_out/m_roi.json 2> gpurun_out/m_roi.err; cat gpurun_out/m_roi.json; tail -5 gpurun_out/m_roi.err
