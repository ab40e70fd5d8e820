set -x
python bench.py --workload tiles1024 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/plain1024.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gather_select|proto_patch|mask_upsample' -s 9 -c 3 -o gpurun_out/prof_r1_gather python bench.py --workload tiles1024 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
