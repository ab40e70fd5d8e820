set -x
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5
python bench.py --workload tiles1024 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['pipeline']['stage_ms'], d['roofline'])"
python bench.py --workload tiles640 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['pipeline']['stage_ms'], d['roofline'])"
