timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02h_pytest.log; cat gpurun_out/r02h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; tail -2 gpurun_out/bench_r02h.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r02h_ref.json 2>> gpurun_out/bench_r02h.err; cut -c1-300 gpurun_out/bench_r02h_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02h_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity --no-backward > gpurun_out/ncu_b.log 2>&1
