python -m pytest tests/test_gpu_slide.py -m gpu -q -x --timeout 600 2>&1 | tail -30
