python -m pytest tests/test_gpu_core.py -x -q -m gpu 2>&1 | tail -2
python tools/nms_phases.py 640 64 1000 1000 2048 2>&1 | tail -8
python tools/nms_phases.py 1024 148 3000 3000 4096 2>&1 | tail -8
