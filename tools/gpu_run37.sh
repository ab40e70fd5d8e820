python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for m in radix bitonic; do
echo "== $m"
HDY_NMS_SORT=$m python tools/nms_phases.py 640 64 1000 1000 2048 2>&1 | tail -8
HDY_NMS_SORT=$m python tools/nms_phases.py 1024 148 3000 3000 4096 2>&1 | tail -8
done
HDY_NMS_SORT=bitonic python -m pytest tests/test_gpu_core.py -x -q -m gpu 2>&1 | tail -2
