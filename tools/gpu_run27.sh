python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload slide --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_slide_n2.json 2> gpurun_out/bench_slide_n2.err; echo rc=$?
tail -5 gpurun_out/bench_slide_n2.err
python bench.py --workload slide --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err; echo rc=$?
tail -3 gpurun_out/bench_slide_n1.err
