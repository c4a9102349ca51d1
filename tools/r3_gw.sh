# tcgen05 gate weight gradient: parity suite, then A/B of the producer pipelining and the mma.sync version
TAG=${1:-r3j}
python -m pytest tests/test_parity_gpu.py -q --timeout 900 -x > gpurun_out/${TAG}_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_parity.log; tail -3 gpurun_out/${TAG}_parity.log
for V in "PTK_GATE_WGRAD_TC=1" "PTK_GATE_WGRAD_TC=0"; do
  env $V python bench.py --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$V.log 2>&1
  python - "$V" "gpurun_out/${TAG}_bench_$V.log" <<PY
import json, sys, re
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        ms = float(re.search(r"\(([\d.]+) ms/step instrumented", r["measured"]).group(1))
        print(sys.argv[1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "gate_wgrad ms", round(r["kernel_time_share"]["gate_wgrad"] * ms, 3), "instrumented", ms)
PY
done
