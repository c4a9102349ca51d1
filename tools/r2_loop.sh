# persistent forward loop kernels (ptk_int_loop.cu / ptk_int_loop2.cu): tests against the chain and the fixtures, then A/B timings
TAG=${1:-r3a}
timeout 1500 python -m pytest tests/test_loop_gpu.py -q --timeout 1400 > gpurun_out/${TAG}_loop_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_loop_test.log
tail -30 gpurun_out/${TAG}_loop_test.log | cut -c1-400
for L in ${LOOP_MODES:-0 2}; do
  PTK_FWD_LOOP=$L timeout 300 python bench.py --eval --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_eval_loop$L.log 2>&1
  PTK_FWD_LOOP=$L timeout 600 python bench.py --no-extra --steps 10 --warmup 3 > gpurun_out/${TAG}_train_loop$L.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_loop?.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "launches", d.get("gpu_launches"), "parity", (d.get("parity") or {}).get("pass"), (d.get("parity") or {}).get("worst_fraction_of_allowed"))
PY
