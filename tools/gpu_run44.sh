python -m pytest tests/test_gpu_slide.py tests/test_gpu_dist.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
python tools/slide_merge_steps.py 100000 2>&1 | tail -1
python __graft_entry__.py smoke 2>&1 | tail -2
python tools/roi_bench.py > gpurun_out/roi_bench.json 2>/dev/null; cat gpurun_out/roi_bench.json
