timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "conv7_operator" > gpurun_out/r2e_conv_test.log 2>&1; echo "conv test rc=$?"; tail -15 gpurun_out/r2e_conv_test.log
timeout 120 python tools/conv_bench.py > gpurun_out/r2e_conv_bench.log 2>&1; tail -8 gpurun_out/r2e_conv_bench.log
