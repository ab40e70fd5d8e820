for e in 0 1; do
for w in tiles640 tiles1024; do
if [ $e = 1 ]; then export HDY_NO_L2_KEEP=1; fi
python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --no-slide --no-e2e | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('nokeep=$e $w', round(d['value']), round(d['ms_per_step'],4), 'one stream', round(d['config']['ms_per_step_one_stream'],4), {k:round(v['ms'],4) for k,v in d['stages'].items()})"
done; done
