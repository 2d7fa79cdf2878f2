timeout 300 python tools/exp_ring2.py 131072 20,22,24,26,28 2>&1 | tail -6
for r in 24 0 24 0; do
IVL_GDN_RING=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity > gpurun_out/bench_r02y_$r.json 2> gpurun_out/bench_r02y.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02y_$r.json'));print('ring $r',d['ms_per_step'],d['kernels']['gdn_layer_ms'],d['e2e']['ms_per_step'],d['clocks']['sm_mhz'])"
done
