python bench.py --workload slide --steps 5 --warmup 2 > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n1.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?
tail -3 gpurun_out/bench_default.err
