M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
for r in 32 24; do
IVL_GDN_RING=$r IVL_GDN_PIPE=1 timeout 300 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r02f_range_ring$r.csv python tools/exp_range.py > gpurun_out/range.log 2>&1
tail -1 gpurun_out/range.log
python - <<P
import csv
rows=[l for l in open('gpurun_out/r02f_range_ring$r.csv') if l.startswith('"')]
for r in csv.DictReader(rows): print('ring $r', r['Metric Name'], r['Metric Value'])
P
done
