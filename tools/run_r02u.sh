python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
for k in gdn_scan_t3_kernel gdn_prep_kernel; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r02e_$k -f python tools/dev_profile.py 131072 2 > gpurun_out/ncu_$k.log 2>&1
ncu -i gpurun_out/r02e_$k.ncu-rep --page raw --csv > gpurun_out/r02e_${k}_raw.csv 2>/dev/null
rm -f gpurun_out/r02e_$k.ncu-rep
tail -2 gpurun_out/ncu_$k.log
done
for f in 0 ""; do
IVL_GDN_FUSED_PREFILL=$f timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-config2 --no-parity > gpurun_out/bench_r02u_f$f.json 2> gpurun_out/bench_r02u.err
python - <<P
import json
d=json.load(open('gpurun_out/bench_r02u_f$f.json'))
c=d['config3_stream']
print('fused="$f"', c['mixers_only'], c['whole_decoder'])
P
done
