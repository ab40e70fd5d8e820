python -m pytest tests/test_gpu_hnet.py -m gpu -q --timeout 900 2>&1 | tail -30
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
