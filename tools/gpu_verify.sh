#!/usr/bin/env bash
# The GPU-side checks and records of a round, in the order they were used (run through gpurun on a B200 box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_verify.sh'
# Everything lands in gpurun_out/; the files quoted in DESIGN.md are copied to profiles/ by hand afterwards.
# A number printed by a run under ncu is never a bench value.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload slide --steps 5 --warmup 3 > gpurun_out/bench_slide_n1.json 2> gpurun_out/bench_slide_n1.err
python bench.py --workload tiles1024 --no-slide --no-cpu-baseline > gpurun_out/bench_tiles1024.json 2> gpurun_out/bench_tiles1024.err
python bench.py --layout 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/bench_layout1.json 2> gpurun_out/bench_layout1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/roi_bench.py > gpurun_out/roi_bench.json 2> gpurun_out/roi_bench.err
# launch list of one tiles640 step sequence (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiles640.csv \
    python bench.py --steps 2 --warmup 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
# one full capture of the dominant kernels (mask phase 1 / 2, filter)
ncu --set full --clock-control none --import-source on -k regex:"mask_upsample_pack|proto_patch|filter_compact_tma" -s 9 -c 3 \
    -o gpurun_out/prof_tiles640 python bench.py --steps 2 --warmup 1 --no-slide --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
