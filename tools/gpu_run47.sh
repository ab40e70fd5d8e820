python -m pytest tests/test_gpu_core.py -x -q -m gpu 2>&1 | tail -3
for f in sparse tma; do
HDY_FILTER=$f python bench.py --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_f_$f.json 2> gpurun_out/bench_f_$f.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_f_$f.json"))
print("$f", d["value"], d["ms_per_step"], d["config"].get("ms_per_step_one_stream"))
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
done
