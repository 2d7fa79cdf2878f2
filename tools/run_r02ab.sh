timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 400 python tools/soak_gdn.py 60 2>&1 | tail -2
