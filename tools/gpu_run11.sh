set -x
python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/plain640.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'proto_patch|mask_upsample_pack|process_mask_kernel|gather_logits|pm_geometry' -s 10 -c 5 -o gpurun_out/prof_r1_mask_v2 python bench.py --workload tiles640 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-slide > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
