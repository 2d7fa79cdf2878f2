for m in overlap split; do
IVL_SHARD_GDN=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | grep '^{' > gpurun_out/bench_r02i_n2_$m.json
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02i_n2_$m.json'))
print('$m', d['ms_per_step'], d['value'], d['dist'], d['kernels'], d['e2e']['value'])
P
done
