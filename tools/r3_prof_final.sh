# final round-2 build: (1) launch list of one full headline step, warm L2 (shares)   (2) --set full of the loop kernels that changed (T = 4)
# (3) --set full of the two all-steps weight-gradient launches at the headline shape
python tools/prof_step.py 64 64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 589 -c 589 --csv --log-file gpurun_out/r3z_launches.csv python tools/prof_step.py 64 64 2 > gpurun_out/prof_ncu1.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r3z_launches.csv > gpurun_out/r3z_launch_list.txt; head -14 gpurun_out/r3z_launch_list.txt
python tools/prof_step.py 4 64 1 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv7_tc2|int_bwd_c_pair|int_bwd_d_pair|int_bwd_a_stream|int_mid|int_post" -s 6 -c 9 -o gpurun_out/r3z_conv python tools/prof_step.py 4 64 1 > gpurun_out/prof_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"wgrad7_ts|gate_wgrad" -c 2 -o gpurun_out/r3z_wgrad python tools/prof_step.py 64 64 1 > gpurun_out/prof_ncu3.log 2>&1
ls -la gpurun_out/r3z_*
