set -x
python -m pytest tests/test_gpu_next.py -x -q -m gpu 2>&1 | tail -15
python tools/roi_bench.py > gpurun_out/roi_bench.json 2> gpurun_out/roi_bench.err; echo rc=$?; cat gpurun_out/roi_bench.json; tail -3 gpurun_out/roi_bench.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:roi_align_levels -c 1 -o gpurun_out/prof_roi python tools/roi_bench.py 8000 256 16 > gpurun_out/ncu_roi.log 2>&1; echo rc=$?
python bench.py --workload slide --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n1.err; cat gpurun_out/bench_slide_n1.json | head -c 3000
