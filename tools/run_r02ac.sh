timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_model_gpu.py tests/test_stream_gpu.py -x -q 2>&1 | tail -4
for pk in 1 0; do
IVL_DECODE_PACKED_PROJ=$pk timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-parity > gpurun_out/bench_r02ac_$pk.json 2> gpurun_out/bench_r02ac.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02ac_$pk.json'))
c=d['config3_stream']
print('packed=$pk', c['mixers_only'], c['whole_decoder'])
P
done
