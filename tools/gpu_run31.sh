ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"merge_|scan_|sort_|keep_keys|overhang|dirty|append_" -c 1200 --csv --log-file gpurun_out/launches_merge.csv python tools/slide_profile.py 100000 > gpurun_out/ncu_merge.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_merge.log
