python -m pytest tests/test_gpu_masks.py -x -q -m gpu 2>&1 | tail -4
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --no-slide --no-cpu-baseline > gpurun_out/bench_trim.json 2> gpurun_out/bench_trim.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_trim.json"))
print(d["value"], d["ms_per_step"], d["config"].get("ms_per_step_one_stream"), d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
