python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload slide --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_slide_n2.json 2> gpurun_out/bench_slide_n2.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_slide_n2.json"))
print(d["value"], d["ms_per_step"], d["slide"]["merge_ms"], d["slide"]["seam_rows"], d["slide"]["exchanges"], d["slide"]["kept"], d["e2e"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu-baseline --no-slide > gpurun_out/bench_default_n2.json 2> gpurun_out/bench_default_n2.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default_n2.json"))
print(d["value"], d["ms_per_step"], d["e2e"])
PY
