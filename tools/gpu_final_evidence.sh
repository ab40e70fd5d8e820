#!/usr/bin/env bash
# Evidence of the round, in the order DESIGN.md quotes it (run through gpurun on a B200 box):
#   gpurun --timeout 1700 -- 'bash tools/gpu_final_evidence.sh'
# A number printed by a run under ncu is never a bench value.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.txt; tail -2 gpurun_out/final_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/final_smoke.txt
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
# launch list of the default command's hot path (our kernels only: the synthetic-input generator launches ~10^5 torch
# kernels before the first timed step); cold-cache, serialised: shares, not absolutes
ncu --kernel-name-base demangled -k regex:hdy:: --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 1500 --csv \
    --log-file gpurun_out/final_launches_slide.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu-baseline --no-torch-cuda --no-e2e > gpurun_out/final_ncu_launches.log 2>&1
# one full capture of the dominant call's kernels inside the slide (traffic of roofline.kernel)
ncu --set full --clock-control none --import-source on -k regex:"proto_patch|mask_upsample_pack2|proto_bin_fused|filter_compact_tma" -s 120 -c 4 \
    -o gpurun_out/final_slide_top python bench.py --steps 2 --warmup 1 --no-sub --no-cpu-baseline --no-torch-cuda --no-e2e > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out/final_*
