timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/exp_p2p.py 2>&1 | grep -v Warning | grep "rank\|Error\|error" | tail -12
bash tools/run_r02k.sh
