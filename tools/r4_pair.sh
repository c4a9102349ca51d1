# two-warps-per-tile reverse-loop kernels: parity with the switch on, then A/B
PTK_BWD_C_PAIR=1 PTK_BWD_D_PAIR=1 python -m pytest tests/test_parity_gpu.py -q --timeout 900 -x > gpurun_out/r4c_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4c_parity.log; tail -3 gpurun_out/r4c_parity.log
for V in "PTK_BWD_C_PAIR=1 PTK_BWD_D_PAIR=1" "PTK_BWD_C_PAIR=1 PTK_BWD_D_PAIR=0" "PTK_BWD_C_PAIR=0 PTK_BWD_D_PAIR=0"; do
  env $V python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith(chr(123)):
        d = json.loads(l); print('$V', round(d['value'], 1), round(d['ms_per_step'], 3), d['roofline']['kernel_time_share']['pointwise_bwd'])
"
done
