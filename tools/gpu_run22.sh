python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload slide --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_slide_n8.json 2> gpurun_out/bench_slide_n8.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline --no-slide > gpurun_out/bench_default_n8.json 2> gpurun_out/bench_default_n8.err; echo rc=$?
tail -3 gpurun_out/bench_default_n8.err
