set -e
CS=infinitevl_b200/csrc
cp infinitevl_b200/lib/libivl_b200.so /tmp/lib_relaxed.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --use_fast_math -Xcompiler -fPIC --expt-relaxed-constexpr -DIVL_BUILDING_DLL -DIVL_NO_RELAXED_WAIT -shared -cudart static -I include -o /tmp/lib_norelax.so $CS/ivl_abi.cu $CS/gdn_prep.cu $CS/gdn_scan.cu $CS/gdn_scan_t.cu $CS/gdn_bwd.cu $CS/gdn_recurrent.cu $CS/gdn_decode.cu $CS/gdn_fused.cu $CS/peer_put.cu $CS/swa_fwd.cu $CS/swa_misc.cu 2>/dev/null
for v in relaxed norelax relaxed norelax; do
cp /tmp/lib_$v.so infinitevl_b200/lib/libivl_b200.so
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));print('$v',d['ms_per_step'],d['kernels'],d['clocks']['sm_mhz'])"
done
cp /tmp/lib_relaxed.so infinitevl_b200/lib/libivl_b200.so
