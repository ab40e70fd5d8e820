#!/usr/bin/env bash
# round 2, third GPU pass: mask kernel paths A/B
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_masks.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2c_pytest.txt
tail -4 gpurun_out/r2c_pytest.txt
for path in 1 2 fused; do
  HDY_MASK_PATH=$path python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 50 > gpurun_out/r2c_t1024_$path.json 2> gpurun_out/r2c_t1024_$path.err
  HDY_MASK_PATH=$path python bench.py --workload tiles640 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 100 > gpurun_out/r2c_t640_$path.json 2> gpurun_out/r2c_t640_$path.err
done
for path in 2 fused; do
HDY_MASK_PATH=$path ncu --set full --clock-control none --import-source on -k regex:"mask_fused|proto_patch|mask_upsample_pack2" -s 6 -c 2 -o gpurun_out/r2c_$path \
    python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 2 --warmup 1 > gpurun_out/r2c_ncu_$path.log 2>&1
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c_pytest_all.txt
tail -3 gpurun_out/r2c_pytest_all.txt
