for k in gdn_scan_t3_kernel gdn_prep_kernel; do
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r02e_$k -f python tools/dev_profile.py 131072 4 > gpurun_out/ncu_$k.log 2>&1
ncu -i gpurun_out/r02e_$k.ncu-rep --page raw --csv > gpurun_out/r02e_${k}_raw.csv 2>/dev/null
rm -f gpurun_out/r02e_$k.ncu-rep
tail -3 gpurun_out/ncu_$k.log
done
