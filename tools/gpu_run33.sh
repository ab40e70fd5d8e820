python tools/slide_merge_steps.py 100000 2>&1 | tail -3
