python tools/prof_step.py 64 64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 589 -c 589 --csv --log-file gpurun_out/r2u_launches.csv python tools/prof_step.py 64 64 2 > gpurun_out/prof_ncu1.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r2u_launches.csv
