set -x
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/pytest_gpu.log
python bench.py --workload tiles1024 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/plain1024.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_tiles1024.csv python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'nms_tiles_kernel|filter_compact_logits' -s 4 -c 4 -o gpurun_out/prof_r1_tiles1024 python bench.py --workload tiles1024 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/plain1024.log | cut -c1-600
