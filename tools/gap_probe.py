"""gap between the last conv CTA finishing and the next (PDL-chained) point-wise kernel getting past its wait."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pathtracker_models_b200 import InT, _lib, synth_pathtracker
torch.manual_seed(0)
m = InT.InT(32, 64, 7).cuda().eval()
x, y = synth_pathtracker(64, 64, 32, 32)
L = _lib.lib()
L.ptk_debug_tc_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.ptk_debug_pw_stamps.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    for it in range(4):
        m(x); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (3 * 128))(); L.ptk_debug_tc_times(buf, 128)
        t = np.array(list(buf), dtype=np.float64).reshape(128, 3)
        pw = (ctypes.c_ulonglong * 4)(); L.ptk_debug_pw_stamps(pw)
        pw = np.array(list(pw), dtype=np.float64)
        print("conv: CTA start spread %.2f us, lifetime med %.2f us, first start -> last end %.2f us | post kernel: entry %+.2f us, wait returned %+.2f us relative to the conv's last CTA end"
              % ((t[:, 0].max() - t[:, 0].min()) / 1e3, np.median(t[:, 2] - t[:, 0]) / 1e3, (t[:, 2].max() - t[:, 0].min()) / 1e3,
                 (pw[0] - t[:, 2].max()) / 1e3, (pw[1] - t[:, 2].max()) / 1e3))
        st = np.sort(t[:, 0] - t[:, 0].min()) / 1e3
        en = np.sort(t[:, 2] - t[:, 0].min()) / 1e3
        md = np.sort(t[:, 1] - t[:, 0].min()) / 1e3
        print("   CTA start offsets us  p0/10/25/50/75/90/100:", np.round(np.percentile(st, [0, 10, 25, 50, 75, 90, 100]), 1))
        print("   CTA mma-done offsets  p0/10/25/50/75/90/100:", np.round(np.percentile(md, [0, 10, 25, 50, 75, 90, 100]), 1))
        print("   CTA end offsets       p0/10/25/50/75/90/100:", np.round(np.percentile(en, [0, 10, 25, 50, 75, 90, 100]), 1))
