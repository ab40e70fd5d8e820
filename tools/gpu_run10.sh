python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8
python bench.py --workload tiles640 --steps 50 --warmup 5 --no-cpu-baseline --no-slide > gpurun_out/bench_tiles640.json 2> gpurun_out/bench_tiles640.err; echo rc=$?; tail -5 gpurun_out/bench_tiles640.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tiles640.json'))
print(d['value'], d['ms_per_step']); print(d['roofline']); print({k:(round(v['ms'],4), v['gbs'] and round(v['gbs'])) for k,v in d['stages'].items()}); print(d['e2e'])"
python bench.py --workload tiles1024 --steps 50 --warmup 5 --no-cpu-baseline --no-slide --no-e2e > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err; echo rc=$?; tail -5 gpurun_out/bench_tiles1024.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tiles1024.json'))
print(d['value'], d['ms_per_step']); print(d['roofline']); print({k:(round(v['ms'],4), v['gbs'] and round(v['gbs'])) for k,v in d['stages'].items()})"

