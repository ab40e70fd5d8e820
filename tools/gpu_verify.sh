#!/usr/bin/env bash
# The GPU-side checks and records of a round (run through gpurun on a B200 box):
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/gpu_verify.sh'
# = tools/gpu_final_evidence.sh (tests, smoke, default bench, reference arm, ncu launch list + one full capture of the
# slide's top kernels) plus the RoIAlign bench.  Multi-GPU: gpurun --gpus N -- 'bash tools/gpu_r2_n8.sh N' (the driver's
# torchrun launch of the default bench).  Everything lands in gpurun_out/; the files quoted in DESIGN.md are copied to
# profiles/ by hand afterwards.  A number printed by a run under ncu is never a bench value.
set -x
bash tools/gpu_final_evidence.sh
python tools/roi_bench.py > gpurun_out/roi_bench.json 2> gpurun_out/roi_bench.err
