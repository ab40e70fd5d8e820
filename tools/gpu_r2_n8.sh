#!/usr/bin/env bash
# the driver's multi-GPU launch of the default bench:  gpurun --gpus N -- 'bash tools/gpu_r2_n8.sh N'
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_bench_default_n$N.json 2> gpurun_out/final_bench_default_n$N.err
tail -c 600 gpurun_out/final_bench_default_n$N.json; tail -3 gpurun_out/final_bench_default_n$N.err
