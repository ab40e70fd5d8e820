set -x
python bench.py --workload tiles1024 --steps 50 --warmup 5 > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err; echo rc=$?
python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain1024.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_tiles1024.csv python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nms_tiles_smem_kernel|filter_compact_tma' -s 6 -c 4 -o gpurun_out/prof_r1_tiles1024_v2 python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
cut -c1-1500 gpurun_out/bench_tiles1024.json
