timeout 600 python -m pytest tests/test_gdn_gpu.py -x -q -k "agree or bit_identical or chunk_matches or varlen or extreme" 2>&1 | tail -5
timeout 300 python tools/dev_tscan.py 2>&1 | tail -8
python tools/trace_tscan.py 3 2>&1 | tail -16
