for l in 1; do
python bench.py --layout $l --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_l$l.json 2> gpurun_out/bench_l$l.err; echo rc=$?; tail -2 gpurun_out/bench_l$l.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_l$l.json"))
print("layout $l", d["value"], d["ms_per_step"], d["config"].get("ms_per_step_one_stream"), d["roofline"]["kernel"], d["roofline"]["frac"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
done
python bench.py --workload tiles1024 --layout 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_1024_l1.json 2> gpurun_out/bench_1024_l1.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_1024_l1.json"))
print("tiles1024 layout 1", d["value"], d["ms_per_step"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
