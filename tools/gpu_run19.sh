for n in 1 2 3 4; do
python bench.py --workload tiles640 --steps 200 --warmup 5 --no-cpu-baseline --no-slide --no-e2e --inflight $n | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('inflight $n', round(d['value']), round(d['ms_per_step'],4), 'one stream', round(d['config']['ms_per_step_one_stream'],4))"
done
python bench.py --workload tiles1024 --steps 100 --warmup 5 --no-cpu-baseline --no-slide --no-e2e --inflight 3 | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('1024 inflight 3', round(d['value']), round(d['ms_per_step'],4), 'one stream', round(d['config']['ms_per_step_one_stream'],4))"
