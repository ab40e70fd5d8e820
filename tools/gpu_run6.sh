set -x
python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain1024.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'filter_compact_tma' -s 6 -c 2 -o gpurun_out/prof_r1_filter_v3 python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
