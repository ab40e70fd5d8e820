python bench.py --no-slide > gpurun_out/bench_default_numa.json 2> gpurun_out/bench_default_numa.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default_numa.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["cpu_baseline"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*at::)" -c 400 --csv --log-file gpurun_out/launches_tiles640.csv python bench.py --steps 2 --warmup 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/ncu_l640.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:nms_tiles_smem -s 3 -c 1 -o gpurun_out/prof_nms_radix python tools/nms_phases.py 1024 148 3000 3000 4096 > gpurun_out/ncu_nms.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:roi_align_levels -c 1 -o gpurun_out/prof_roi2 python tools/roi_bench.py 8000 256 16 > gpurun_out/ncu_roi2.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"merge_round_kernel|merge_large_push" -c 3 -o gpurun_out/prof_merge_round python tools/slide_merge_steps.py 40000 > gpurun_out/ncu_mr.log 2>&1; echo rc=$?
