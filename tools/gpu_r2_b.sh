#!/usr/bin/env bash
# round 2, second GPU pass: fused mask kernel parity + A/B, slide bench
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2b_pytest.txt
tail -4 gpurun_out/r2b_pytest.txt
python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 50 > gpurun_out/r2b_t1024_fused.json 2> gpurun_out/r2b_t1024_fused.err
HDY_MASK_PATH=2phase python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 50 > gpurun_out/r2b_t1024_2phase.json 2> gpurun_out/r2b_t1024_2phase.err
python bench.py --workload tiles640 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 100 > gpurun_out/r2b_t640_fused.json 2> gpurun_out/r2b_t640_fused.err
python bench.py --steps 3 --warmup 3 --no-sub --no-e2e --no-cpu-baseline > gpurun_out/r2b_slide.json 2> gpurun_out/r2b_slide.err
tail -3 gpurun_out/r2b_slide.err
ncu --set full --clock-control none --import-source on -k regex:"mask_fused" -s 4 -c 1 -o gpurun_out/r2b_fused \
    python bench.py --workload tiles1024 --no-sub --no-cpu-baseline --no-e2e --no-torch-cuda --steps 2 --warmup 1 > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
