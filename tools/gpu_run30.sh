python tools/slide_profile.py 100000 2>&1 | tail -30
python tools/roi_bench.py 2>&1 | tail -2
python -m pytest tests/test_gpu_next.py -x -q -m gpu -k roi 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_slide.csv python tools/slide_profile.py 100000 > gpurun_out/ncu_slide.log 2>&1; echo rc=$?
