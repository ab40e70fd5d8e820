python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?; tail -2 gpurun_out/bench_default.err
python bench.py --workload slide --steps 5 --warmup 3 > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?; tail -2 gpurun_out/bench_slide_n1.err
python bench.py --workload tiles1024 --no-slide --no-cpu-baseline > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err; echo rc=$?
python - <<PY
import json
for f in ("bench_default","bench_slide_n1","bench_tiles1024"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d.get("e2e"), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
    print("   ", {k: round(v["ms"],4) for k,v in d["stages"].items()})
    if "slide" in d: print("    slide", d["slide"]["ms_per_slide"], d["slide"]["merge_ms"])
PY
