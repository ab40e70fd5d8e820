set -x
python bench.py --workload tiles640 --masks paste --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-slide --inflight 1 > gpurun_out/plain640.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'paste_masks_kernel' -s 8 -c 1 -o gpurun_out/prof_r1_paste python bench.py --workload tiles640 --masks paste --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-slide --inflight 1 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
