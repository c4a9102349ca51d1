# persistent forward loop (ptk_int_loop.cu): bit-identity test against the chain, then A/B timings (eval and training)
TAG=${1:-r3a}
timeout 900 python -m pytest tests/test_loop_gpu.py -x -q --timeout 800 > gpurun_out/${TAG}_loop_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_loop_test.log
tail -25 gpurun_out/${TAG}_loop_test.log
grep -q "rc=0" gpurun_out/${TAG}_loop_test.log || exit 1
for L in 0 1; do
  PTK_FWD_LOOP=$L timeout 300 python bench.py --eval --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_eval_loop$L.log 2>&1
  PTK_FWD_LOOP=$L timeout 600 python bench.py --no-extra --steps 10 --warmup 3 > gpurun_out/${TAG}_train_loop$L.log 2>&1
done
python - <<PY
import json
for f in ("eval_loop0", "eval_loop1", "train_loop0", "train_loop1"):
    for l in open("gpurun_out/${TAG}_%s.log" % f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "launches", d.get("gpu_launches"), "parity", (d.get("parity") or {}).get("pass"))
PY
