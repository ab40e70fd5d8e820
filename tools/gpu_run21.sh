set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --workload slide --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_slide_n4.json 2> gpurun_out/bench_slide_n4.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default_n4.json 2> gpurun_out/bench_default_n4.err; echo rc=$?
tail -3 gpurun_out/bench_default_n4.err
python -m pytest tests/test_gpu_masks.py -m gpu -q -x 2>&1 | tail -2
