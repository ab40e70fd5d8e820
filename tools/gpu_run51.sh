python -m pytest tests/test_gpu_core.py -x -q -m gpu 2>&1 | tail -3
python bench.py --layout 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_l1.json 2> gpurun_out/bench_l1.err; echo rc=$?; tail -2 gpurun_out/bench_l1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_l1.json"))
print("layout 1", d["value"], d["ms_per_step"], d["config"].get("ms_per_step_one_stream"), d["roofline"]["kernel"], d["roofline"]["frac"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
