"""per-CTA globaltimer stamps of one tcgen05 conv launch: launch ramp, CTA lifetime spread, tail."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pathtracker_models_b200 import _lib
from pathtracker_models_b200.functional import conv7
torch.manual_seed(0)
x = torch.randn(64, 32, 32, 32, device="cuda"); w = torch.randn(32, 32, 7, 7, device="cuda") * 0.05
L = _lib.lib()
L.ptk_debug_tc_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); y = conv7(x, w, conv_mode=1); e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (3 * 128))()
    L.ptk_debug_tc_times(buf, 128)
    t = np.array(list(buf), dtype=np.float64).reshape(128, 3)
    t0 = t[:, 0].min()
    print("event %.1f us | CTA start spread %.1f us | lifetime min/med/max %.1f/%.1f/%.1f us | first start -> last end %.1f us | mma-done med %.1f us"
          % (e0.elapsed_time(e1) * 1e3, (t[:, 0].max() - t0) / 1e3, *(np.percentile(t[:, 2] - t[:, 0], [0, 50, 100]) / 1e3),
             (t[:, 2].max() - t0) / 1e3, np.median(t[:, 1] - t[:, 0]) / 1e3))
