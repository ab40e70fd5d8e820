python -m pytest tests/test_gpu_masks.py -m gpu -q -x 2>&1 | tail -2
for w in tiles640 tiles1024; do
python bench.py --workload $w --masks paste --steps 50 --warmup 5 --no-slide --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
L=[l for l in sys.stdin]
J=[l for l in L if l.startswith('{')]
if not J: print(''.join(L[-15:])); sys.exit()
d=json.loads(J[-1]); print('$w', round(d['value']), round(d['ms_per_step'],4), 'one stream', round(d['config']['ms_per_step_one_stream'],4), {k:round(v['ms'],4) for k,v in d['stages'].items()}, d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done
