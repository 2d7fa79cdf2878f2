IVL_BENCH_LAYER_TRACE=2 IVL_SHARD_TRANSPORT=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2> gpurun_out/bench_r02n.err | grep '^{' > gpurun_out/bench_r02n.json
grep -A7 "layer-trace" gpurun_out/bench_r02n.err
