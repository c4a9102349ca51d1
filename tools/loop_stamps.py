"""per-phase cycle sums of the persistent forward loop kernel (PTK_DEBUG_STAMPS=1 build, PTK_FWD_LOOP=1): thread 0 of every CTA."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pathtracker_models_b200 import InT, _lib, synth_pathtracker
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
train = len(sys.argv) > 2 and sys.argv[2] == "train"
torch.manual_seed(0)
m = InT.InT(32, T, 7).cuda()
if not train: m.eval()
x, y = synth_pathtracker(64, T, 32, 32)
L = _lib.lib()
L.ptk_debug_loop_stamps.argtypes = [ctypes.c_void_p]
names = ["conv0 (thread 0 = producer warp)", "conv0 rest of the CTA (epilogue)", "mid head (warp 0)", "BN0 barrier wait", "BN0 fold", "mid tail (warp 0)",
         "sample barrier 0", "conv1 (thread 0)", "conv1 rest of the CTA", "post loads", "BN1 barrier wait", "BN1 fold", "post tile (warp 0)", "sample barrier 1"]
for it in range(3):
    if train:
        out, _ = m(x); out.sum().backward()
    else:
        with torch.no_grad(): m(x)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (148 * 16))(); L.ptk_debug_loop_stamps(buf)
a = np.array(list(buf), dtype=np.float64).reshape(148, 16)[:128, :14] / T / 1965.0   # us per frame at 1965 MHz
print("per frame, us (mean / min / max over the 128 CTAs); total %.1f us per frame" % a.sum(1).mean())
for i, n in enumerate(names):
    print("  %-36s %6.2f %6.2f %6.2f" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))
