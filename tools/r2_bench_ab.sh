# A/B: bench.py (headline only) with the second- and first-generation conv kernels
for v in 0 1; do
PTK_CONV_V1=$v python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ab_v1_$v.log 2>&1
python - <<PY
import json
for l in open("gpurun_out/ab_v1_$v.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("PTK_CONV_V1=$v value %.1f ms %.3f e2e %.1f conv launch %.2f us" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"]*1e3))
        print(d["roofline"]["kernel_time_share"], d["roofline"]["measured"])
PY
done
