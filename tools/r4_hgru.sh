# hGRU with the staged reverse-loop kernels: parity (all hGRU cases + the rest), then FFhGRU A/B
python -m pytest tests/test_parity_gpu.py -q --timeout 900 -x > gpurun_out/r4b_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4b_parity.log; tail -3 gpurun_out/r4b_parity.log
for V in "PTK_X=1" "PTK_BWD_C_STAGED=0 PTK_BWD_D_STAGED=0"; do
  env $V python bench.py --model hgru --no-extra --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith(chr(123)):
        d = json.loads(l); print('$V', round(d['value'], 1), round(d['ms_per_step'], 3))
"
done
