#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_kernel<\(int\)14, \(bool\)1>' -s 3 -c 1 -o gpurun_out/zz_roi_cl python tools/roi_bench.py 8192 > gpurun_out/zz_ncu.log 2>&1
tail -2 gpurun_out/zz_ncu.log
