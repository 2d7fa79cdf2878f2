python tools/trace_tscan.py 1 2>&1 | tail -20
IVL_NVCC_EXTRA="-DIVL_TSCAN_NA=3" python tools/trace_tscan.py 1 2>&1 | tail -20
python tools/trace_scan.py 2>&1 | tail -4
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_elapsed
IVL_GDN_PIPE=1 timeout 300 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r02d_range_overlapped.csv python tools/exp_range.py > gpurun_out/range.log 2>&1
tail -2 gpurun_out/range.log
