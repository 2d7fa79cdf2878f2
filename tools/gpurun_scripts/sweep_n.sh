N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --sweep --no-gpu-reference 2> gpurun_out/sweep_n$N.err | grep '^{' | cut -c1-200
grep -i "unavailable\|Traceback" gpurun_out/sweep_n$N.err | head -3
