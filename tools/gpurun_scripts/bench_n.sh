N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r02_n$N.err | grep '^{' > gpurun_out/bench_r02_n$N.json
grep -i "unavailable\|Error" gpurun_out/bench_r02_n$N.err | grep -v Warning | head -5
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02_n$N.json'))
print('N=$N', d['ms_per_step'], d['value'], {k:v for k,v in d['dist'].items() if k!='parity_err'}, d['e2e']['value'], d['clocks'])
print(d['dist'].get('parity_err',{}).get('operators'), d['dist'].get('parity_err',{}).get('decoder'))
P
