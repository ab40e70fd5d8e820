python -m pytest tests/test_gpu_slide.py tests/test_gpu_dist.py tests/test_gpu_pipeline.py tests/test_gpu_hnet.py -x -q -m gpu 2>&1 | tail -3
python tools/slide_profile.py 100000 2>&1 | tail -14
