"""micro-benchmark of the 7x7 convolution operator (ptk_conv7) in both modes; CUDA-event timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathtracker_models_b200.functional import conv7

B, H, W = int(os.environ.get("B", 64)), 32, 32
iters = int(os.environ.get("ITERS", 20))
x = torch.randn(B, 32, H, W, device="cuda")
w = torch.randn(32, 32, 7, 7, device="cuda") * 0.05
flop = 2 * B * H * W * 32 * 32 * 49
for mode, name in ((0, "fp32-ffma"), (1, "bf16x3-tcgen05")):
    for acc in (False, True):
        out = torch.zeros_like(x) if acc else None
        for _ in range(3):
            conv7(x, w, conv_mode=mode, out=out, want_stats=not acc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            conv7(x, w, conv_mode=mode, out=out, want_stats=not acc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{name:16s} accumulate={acc!s:5s} {ms*1e3:8.1f} us/call (incl. weight prep launch)  {flop/ms/1e9:8.1f} TFLOP/s")

import ctypes
from pathtracker_models_b200 import _lib
L = _lib.lib()
buf = (ctypes.c_longlong * 8)()
conv7(x, w, conv_mode=1, want_stats=True); torch.cuda.synchronize()
L.ptk_debug_tc_stamps(buf)
t = list(buf)
print("tc conv CTA(0) cycles: setup+zero %d, staging %d, mma %d, epilogue %d, total %d" % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[4]-t[0]))
out = torch.zeros_like(x)
conv7(x, w, conv_mode=1, out=out); torch.cuda.synchronize()
L.ptk_debug_tc_stamps(buf)
t = list(buf)
print("tc conv (accumulate) cycles: setup+zero %d, staging %d, mma %d, epilogue %d, total %d" % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[4]-t[0]))

import numpy as np
for acc in (False,):
    conv7(x, w, conv_mode=1, want_stats=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); conv7(x, w, conv_mode=1, want_stats=True); e1.record(); torch.cuda.synchronize()
    n = B * 2
    tb = (ctypes.c_ulonglong * (3 * n))()
    L.ptk_debug_tc_times(tb, n)
    t = np.array(list(tb), dtype=np.float64).reshape(n, 3)
    t0 = t[:, 0].min()
    print("event time of the call (prep + memset + conv): %.1f us" % (e0.elapsed_time(e1) * 1e3))
    print("CTA start skew (us): min %.1f  median %.1f  max %.1f" % (0, np.median(t[:, 0] - t0) / 1e3, (t[:, 0].max() - t0) / 1e3))
    print("CTA lifetime (us): min %.1f median %.1f max %.1f" % tuple(np.percentile(t[:, 2] - t[:, 0], [0, 50, 100]) / 1e3))
    print("CTA mma-done - start (us): median %.1f max %.1f" % tuple(np.percentile(t[:, 1] - t[:, 0], [50, 100]) / 1e3))
    print("kernel span first start -> last end: %.1f us" % ((t[:, 2].max() - t0) / 1e3))
