timeout 900 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -15
for m in p2p nccl; do
IVL_SHARD_TRANSPORT=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02k_n2_$m.err | grep '^{' > gpurun_out/bench_r02k_n2_$m.json
grep -v Warning gpurun_out/bench_r02k_n2_$m.err | grep -i "error\|unavailable\|Traceback" | head -5
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02k_n2_$m.json'))
print('$m', d['ms_per_step'], d['value'], {k:v for k,v in d['dist'].items() if k!='parity_err'}, d['dist'].get('parity_err',{}).get('operators'), d['dist'].get('parity_err',{}).get('decoder'), d['e2e']['value'])
P
done
