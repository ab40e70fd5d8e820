python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 2>&1 | tail -15
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-slide | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), d['roofline'], d['cpu_baseline'])"
