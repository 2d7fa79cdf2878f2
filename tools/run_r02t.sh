timeout 900 python -m pytest tests/test_swa_gpu.py tests/test_modules_gpu.py -x -q 2>&1 | tail -12
