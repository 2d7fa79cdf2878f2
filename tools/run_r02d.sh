M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_elapsed
IVL_GDN_PIPE=1 timeout 300 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r02d_range_overlapped.csv python tools/exp_range.py > gpurun_out/range.log 2>&1
tail -2 gpurun_out/range.log
IVL_SWA_POLY=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity > gpurun_out/bench_r02d_poly3.json 2> gpurun_out/bench_r02d_poly3.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02d_poly3.json'));print('poly3',d['ms_per_step'],d['kernels'],d['clocks'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity > gpurun_out/bench_r02d_poly0.json 2> gpurun_out/bench_r02d_poly0.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r02d_poly0.json'));print('poly0',d['ms_per_step'],d['kernels'],d['clocks'])"
timeout 900 python bench.py --sweep > gpurun_out/sweep_n1.log 2>&1
tail -12 gpurun_out/sweep_n1.log | cut -c1-400
