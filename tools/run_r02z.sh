timeout 400 python tools/soak_gdn.py 150 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gdn_gpu.py tests/test_stream_gpu.py -x -q 2>&1 | tail -3
