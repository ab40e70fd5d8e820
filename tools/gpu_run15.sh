for v in 24_24 48_12 32_18 64_9 96_6; do
cp tools/micro/libs/lib_$v.so hd_yolo_b200/libhdyolo_b200.so
echo "== region $v"
python -m pytest tests/test_gpu_masks.py -m gpu -q -x -k process_mask 2>&1 | tail -1
for w in tiles640 tiles1024; do
python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-slide --no-e2e | python -c "
import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$w', round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['stages'].items() if 'mask' in k})"
done
done
