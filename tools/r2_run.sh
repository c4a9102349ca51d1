# usage: bash tools/r2_run.sh TAG   -- GPU tests + default bench, logs under gpurun_out/
TAG=${1:-r2x}
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -15 gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
        print(d["roofline"]["kernel_time_share"], d["roofline"]["launch_ms"])
PY
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-600
