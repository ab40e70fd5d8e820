python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
python bench.py --workload tiles640 --masks paste --steps 50 --warmup 5 --no-slide --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('paste640', round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['stages'].items()})"
for w in tiles640 tiles1024; do
python bench.py --workload $w --steps 100 --warmup 5 --no-slide --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$w', round(d['value']), round(d['ms_per_step'],4), 'one stream', round(d['config']['ms_per_step_one_stream'],4), {k:round(v['ms'],4) for k,v in d['stages'].items()})"
done
