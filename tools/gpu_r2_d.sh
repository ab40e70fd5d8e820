#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2d_pytest.txt
tail -4 gpurun_out/r2d_pytest.txt
for path in 1 2 fused; do
  HDY_MASK_PATH=$path python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 50 > gpurun_out/r2d_t1024_$path.json 2> gpurun_out/r2d_t1024_$path.err
done
python bench.py --workload tiles640 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 100 > gpurun_out/r2d_t640.json 2> gpurun_out/r2d_t640.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2d_slide.json 2> gpurun_out/r2d_slide.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda --slide-streams 1 > gpurun_out/r2d_slide_s1.json 2> gpurun_out/r2d_slide_s1.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda --dtype f16 > gpurun_out/r2d_slide_f16.json 2> gpurun_out/r2d_slide_f16.err
tail -3 gpurun_out/r2d_slide*.err
ncu --set full --clock-control none --import-source on -k regex:"proto_patch|mask_upsample_pack2" -s 6 -c 2 -o gpurun_out/r2d_p2 \
    python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 2 --warmup 1 > gpurun_out/r2d_ncu.log 2>&1
