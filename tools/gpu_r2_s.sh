#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.txt; cat gpurun_out/final_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; tail -2 gpurun_out/final_bench_default.err
