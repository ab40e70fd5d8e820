python -m pytest tests/test_gpu_next.py -x -q -m gpu 2>&1 | tail -2
python tools/roi_bench.py 2>/dev/null | tail -1
