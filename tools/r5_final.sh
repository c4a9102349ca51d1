# final build of round 2 (after the ConvLSTM in-graph penalty): GPU suite, default bench, then the launch list of one headline step (warm L2)
# and --set full of the loop kernels (T = 4) -> gpurun_out/r5z_*
TAG=r5z
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"
python tools/prof_step.py 64 64 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 589 -c 589 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_step.py 64 64 2 > gpurun_out/prof_ncu1.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_list.txt; head -14 gpurun_out/${TAG}_launch_list.txt
python tools/prof_step.py 4 64 1 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv7_tc2|int_bwd_c_pair|int_bwd_d_pair|int_bwd_a_stream|int_mid|int_post" -s 6 -c 9 -o gpurun_out/${TAG}_loop python tools/prof_step.py 4 64 1 > gpurun_out/prof_ncu2.log 2>&1
ls -la gpurun_out/${TAG}_*
