for ns in 1 2 3; do
python bench.py --workload slide --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --slide-streams $ns > gpurun_out/bench_slide_s$ns.json 2> gpurun_out/bench_slide_s$ns.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_slide_s$ns.json"))
print("streams $ns", d["ms_per_step"], d["slide"]["merge_ms"], d["slide"]["kept"], d["slide"]["detections"])
PY
done
