#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_masks.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2e_pytest.txt
tail -3 gpurun_out/r2e_pytest.txt
python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 50 > gpurun_out/r2e_t1024.json 2> gpurun_out/r2e_t1024.err
python bench.py --workload tiles640 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 100 > gpurun_out/r2e_t640.json 2> gpurun_out/r2e_t640.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline --no-torch-cuda > gpurun_out/r2e_slide.json 2> gpurun_out/r2e_slide.err
ncu --set full --clock-control none --import-source on -k regex:"mask_upsample_pack2" -s 3 -c 1 -o gpurun_out/r2e_p2 \
    python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 2 --warmup 1 > gpurun_out/r2e_ncu.log 2>&1
