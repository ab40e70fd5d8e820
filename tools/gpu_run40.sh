python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/nms_phases.py 1024 148 3000 3000 4096 2>&1 | tail -8
python tools/nms_phases.py 640 64 1000 1000 2048 2>&1 | tail -8
python tools/slide_merge_steps.py 100000 2>&1 | tail -1
