python tools/prof_step.py 64 64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 589 -c 589 --csv --log-file gpurun_out/r2z_launches.csv python tools/prof_step.py 64 64 2 > gpurun_out/prof_ncu1.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2z_launches.csv | head -12
python tools/prof_step.py 4 64 1 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv7_tc2|wgrad7_ts|int_bwd_d_staged|int_bwd_c_staged|int_bwd_a_stream|gate_wgrad|int_mid|int_post" -s 14 -c 26 -o gpurun_out/r2z_full python tools/prof_step.py 4 64 1 > gpurun_out/prof_ncu2.log 2>&1
ls -la gpurun_out/r2z_*
