set -x
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
python bench.py --workload tiles1024 --steps 100 --warmup 5 --no-slide > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err; echo rc=$?
python bench.py --masks paste --steps 100 --warmup 5 --no-slide > gpurun_out/bench_paste640.json 2> gpurun_out/bench_paste640.err; echo rc=$?
python bench.py --workload slide --steps 5 --warmup 2 > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?
python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide --inflight 1 > gpurun_out/plain640.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1_tiles640.csv python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide --inflight 1 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'proto_patch|mask_upsample|filter_compact_tma|nms_tiles_smem|gather_select' -s 14 -c 5 -o gpurun_out/prof_r1_final640 python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide --inflight 1 > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
