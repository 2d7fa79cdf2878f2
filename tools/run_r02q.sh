for mc in 32 8; do
CUDA_DEVICE_MAX_CONNECTIONS=$mc IVL_BENCH_LAYER_TRACE=0 IVL_SHARD_TRANSPORT=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2> gpurun_out/bench_r02p.err | grep '^{' > gpurun_out/bench_r02p.json
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02p.json'))
print('p2p maxconn=$mc', d['ms_per_step'], d['dist']['single_prompt_ms'], d['clocks']['sm_mhz'])
P
done
