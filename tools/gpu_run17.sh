for v in 4,2,9216 4,2,10496 5,2,10496 8,2,5248 11,2,5248 4,3,5248 4,2,15744 2,2,20992 7,3,5248; do
echo "variant $v"
HDY_TMA_VARIANT=$v python bench.py --workload tiles640 --steps 100 --warmup 5 --no-cpu-baseline --no-slide --no-e2e | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],4), round(d['stages']['hdy_filter_compact_logits']['ms'],4), round(d['stages']['hdy_filter_compact_logits']['gbs']))"
done
