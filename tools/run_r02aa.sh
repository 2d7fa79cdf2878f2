IVL_GDN_PREP_TC=1 timeout 600 python -m pytest tests/test_gdn_gpu.py -x -q -k "chunk_matches or extreme or no_l2norm or varlen or agree or triton_fixture" 2>&1 | tail -6
IVL_GDN_PREP_TC=1 timeout 300 python tools/dev_tscan.py 2>&1 | grep "T=131072 tscan=3\|T=4160 H=4 h0=f32 tscan=3"
IVL_GDN_PREP_TC=0 timeout 300 python tools/dev_tscan.py 2>&1 | grep "T=131072 tscan=3"
