python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
for v in 4,2,9216 5,2,9216 11,2,9216 8,2,4608 16,2,4608 7,3,9216 11,2,4608; do
echo "variant $v"
HDY_TMA_VARIANT=$v python bench.py --workload tiles1024 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['pipeline']['stage_ms']['hdy_filter_compact_logits'], d['roofline']['frac'])"
done
HDY_TMA_VARIANT=11,2,9216 python bench.py --workload tiles640 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['pipeline']['stage_ms'], d['roofline']['frac'])"
