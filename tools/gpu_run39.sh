python -m pytest tests/test_gpu_core.py tests/test_gpu_slide.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
python tools/nms_phases.py 1024 148 3000 3000 4096 2>&1 | tail -8
python tools/nms_phases.py 640 64 1800 2000 2048 2>&1 | tail -8
python bench.py --workload slide --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --slide-streams 3 > gpurun_out/bench_slide_s3.json 2> gpurun_out/bench_slide_s3.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_slide_s3.json"))
print("slide streams 3", d["ms_per_step"], d["slide"]["merge_ms"], d["slide"]["kept"], d["slide"]["detections"])
print({k: round(v["ms"],4) for k,v in d["stages"].items()})
PY
