#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python tools/mask_batch_probe.py 3328 > gpurun_out/r2f_probe.txt 2>&1
python tools/mask_batch_probe.py 3072 >> gpurun_out/r2f_probe.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2f_probe_ncu.csv python tools/mask_batch_probe.py 3328 > /dev/null 2>&1
cat gpurun_out/r2f_probe.txt
