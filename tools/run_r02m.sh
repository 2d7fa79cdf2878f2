for m in p2p; do
IVL_SHARD_TRANSPORT=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2> gpurun_out/bench_r02m_n2_$m.err | grep '^{' > gpurun_out/bench_r02m_n2_$m.json
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02m_n2_$m.json'))
print('$m', d['ms_per_step'], d['value'], {k:v for k,v in d['dist'].items() if k!='parity_err'}, d['e2e']['value'], d['clocks'])
P
done
