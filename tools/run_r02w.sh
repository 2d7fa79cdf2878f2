timeout 300 python tools/exp_ring2.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gdn_gpu.py -x -q -k "bit_identical or varlen or graph" 2>&1 | tail -4
