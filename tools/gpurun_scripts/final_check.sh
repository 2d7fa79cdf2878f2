timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:swa_fwd_kernel -s 2 -c 1 -o gpurun_out/r02i_swa -f python tools/dev_profile.py 131072 4 > gpurun_out/ncu_swa.log 2>&1
ncu -i gpurun_out/r02i_swa.ncu-rep --page raw --csv > gpurun_out/r02i_swa_raw.csv 2>/dev/null
rm -f gpurun_out/r02i_swa.ncu-rep; tail -1 gpurun_out/ncu_swa.log
