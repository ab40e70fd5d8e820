#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2j_smoke.txt 2>&1; tail -3 gpurun_out/r2j_smoke.txt
SECONDS=0
python bench.py > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err
echo "default bench took $SECONDS s" | tee gpurun_out/r2j_time.txt
tail -3 gpurun_out/r2j_bench_default.err
SECONDS=0
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err
echo "reference arm took $SECONDS s" | tee -a gpurun_out/r2j_time.txt
