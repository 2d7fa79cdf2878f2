timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for m in 3 1; do
IVL_GDN_TSCAN=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity > gpurun_out/bench_r02h_t$m.json 2> gpurun_out/bench_r02h_t$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02h_t$m.json'));print('tscan$m',d['ms_per_step'],d['kernels'],d['clocks']['sm_mhz'])"
done
