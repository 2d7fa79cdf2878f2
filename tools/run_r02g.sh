timeout 900 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02g_n2.json 2> gpurun_out/bench_r02g_n2.err
tail -3 gpurun_out/bench_r02g_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r02g_n2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','dist','e2e')})
P
