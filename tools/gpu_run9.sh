python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5
python tools/slide_profile.py 30000
python tools/slide_profile.py 100000 2>&1 | tail -14
