M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
IVL_GDN_RING=28 IVL_GDN_PIPE=1 timeout 200 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file gpurun_out/r02f_range_ring28.csv python tools/exp_range.py > gpurun_out/range.log 2>&1
python - <<P
import csv
rows=[l for l in open('gpurun_out/r02f_range_ring28.csv') if l.startswith('"')]
for r in csv.DictReader(rows): print('ring 28', r['Metric Name'], r['Metric Value'])
P
for r in 28 24 28 24; do
IVL_GDN_RING=$r timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-config3 --no-parity --no-backward > gpurun_out/bench_ring.json 2> gpurun_out/bench_ring.err
python -c "
import json;d=json.load(open('gpurun_out/bench_ring.json'));print('ring $r',d['ms_per_step'],d['kernels']['gdn_layer_ms'],d['clocks']['sm_mhz'])"
done
