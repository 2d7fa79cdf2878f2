IVL_BENCH_NOCOMM=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | grep '^{' > gpurun_out/bench_r02j_n2_nocomm.json
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02j_n2_nocomm.json'))
print('nocomm', d['ms_per_step'], d['value'], d['dist'], d['kernels'], d['clocks'])
P
