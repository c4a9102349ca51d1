"""print per-tensor parity errors (vs the fp64 reference fixtures) of the CUDA path, both conv modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import load_case, rel_err
from tests.test_parity_gpu import _model, _run, CASES, LONG_CASES
from pathtracker_models_b200 import _lib

for name, kind, kw in CASES + LONG_CASES:
    case = load_case(name)
    for mode, mname in ((0, "fp32"), (1, "bf16x3"), (2, "bf16")):
        m = _model(case, kind, **kw)
        m.conv_mode = mode
        out, loss, grads, _ = _run(case, m)
        errs = [("logits", rel_err(out, case["logits"]), case["noise"].get("__logits__", 0))]
        for n, g in case["grads"].items():
            errs.append((n, rel_err(grads[n], g), case["noise"].get(n, 0)))
        worst = max(errs, key=lambda e: e[1])
        ratio = max(errs, key=lambda e: e[1] / max(1e-4, 4 * e[2]))
        print(f"{name:18s} {mname:7s} worst {worst[0]:28s} {worst[1]:.2e} (ref fp32 noise {worst[2]:.2e})   "
              f"tightest vs gate: {ratio[0]:24s} {ratio[1] / max(1e-4, 4 * ratio[2]):.2f} of allowed")
