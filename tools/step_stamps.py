"""run one small fwd+bwd in tensor-core mode and print the in-kernel cycle stamps of the TC kernels."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathtracker_models_b200 import InT, _lib, synth_pathtracker
torch.manual_seed(0)
m = InT.InT(32, 8, 7).cuda(); m.conv_mode = 1
x, y = synth_pathtracker(64, 4, 32, 32)
for _ in range(2):
    out, _ = m(x); out.sum().backward()
torch.cuda.synchronize()
L = _lib.lib(); buf = (ctypes.c_longlong * 8)()
L.ptk_debug_wt_stamps(buf); t = list(buf)
print("wgrad_tc CTA0 cycles: setup+zero %d, staging %d, mma-issue-loop %d, total(epilogue end) %d" % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[0]))
L.ptk_debug_tc_stamps(buf); t = list(buf)
print("conv_tc  CTA0 cycles: setup+zero %d, staging %d, mma %d, epilogue %d, total %d" % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[4]-t[0]))
