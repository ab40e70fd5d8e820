set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
tail -5 gpurun_out/bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json'))
print(d['value'], d['ms_per_step'], d['n_gpus']); print(d['slide']); print(d['e2e']); print(d['cpu_baseline'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload slide --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_slide_n2.json 2> gpurun_out/bench_slide_n2.err; echo rc=$?
tail -5 gpurun_out/bench_slide_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_slide_n2.json'))
print(d['value'], d['ms_per_step'], d['n_gpus']); print(d['slide']); print(d['e2e'])"
