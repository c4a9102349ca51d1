# usage: bash tools/r2_ab_env.sh VAR  -- headline bench with VAR unset and VAR=1
for v in "" 1; do
env $1=$v python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ab_env_$v.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/ab_env_$v.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$1=$v value %.1f ms %.3f launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]), d["roofline"]["kernel_time_share"])
PY
done
