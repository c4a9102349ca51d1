# launch list of one full headline step (shares) + full captures of the hot kernels (short horizon: same kernels per step)
python tools/prof_step.py 64 64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 717 -c 717 --csv --log-file gpurun_out/r2i_launches.csv python tools/prof_step.py 64 64 2 > gpurun_out/prof_ncu1.log 2>&1
python tools/prof_step.py 6 64 1 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv7_tc2|wgrad7_ts|int_bwd_d|int_bwd_c|gate_wgrad|int_mid|int_post" -s 20 -c 16 -o gpurun_out/r2i_full python tools/prof_step.py 6 64 1 > gpurun_out/prof_ncu2.log 2>&1
ls -la gpurun_out/r2i_*
