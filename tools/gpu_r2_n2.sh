#!/usr/bin/env bash
# 2-GPU pass: NCCL slide (digest must equal the 1-GPU run), per-phase profile of the sharded merge
set -x
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slide_dist_profile.py 100000 1 > gpurun_out/r2_dist_profile_n$N.txt 2> gpurun_out/r2_dist_profile_n$N.err
tail -3 gpurun_out/r2_dist_profile_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-sub --no-cpu-baseline --no-torch-cuda > gpurun_out/r2_slide_n$N.json 2> gpurun_out/r2_slide_n$N.err
tail -2 gpurun_out/r2_slide_n$N.err
