python -m pytest tests/test_gpu_slide.py tests/test_gpu_pipeline.py tests/test_gpu_dist.py -m gpu -q -x --timeout 900 2>&1 | tail -25
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
python tools/slide_profile.py 100000 2>&1 | tail -12
