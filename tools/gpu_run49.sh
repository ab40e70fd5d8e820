python -m pytest tests/test_gpu_core.py -x -q -m gpu 2>&1 | tail -3
for l in 1 0; do
python bench.py --layout $l --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_l$l.json 2> gpurun_out/bench_l$l.err; echo rc=$?; tail -2 gpurun_out/bench_l$l.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_l$l.json"))
print("layout $l", d["value"], d["ms_per_step"], d["config"].get("ms_per_step_one_stream"), d["roofline"]["kernel"], d["roofline"]["frac"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
done
