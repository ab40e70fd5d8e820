set -x
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
tail -3 gpurun_out/bench_default.err
python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/plain640.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1_tiles640.csv python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/ncu_l.log 2>&1
python bench.py --workload tiles1024 --steps 100 --warmup 5 > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err; echo rc=$?
python bench.py --workload slide --steps 5 --warmup 2 > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?
