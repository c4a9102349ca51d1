"""one training step (fwd + bwd) of the headline shape for ncu: python tools/prof_step.py [T] [B] [steps] [size]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pathtracker_models_b200 import InT, synth_pathtracker
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
HW = int(sys.argv[4]) if len(sys.argv) > 4 else 32
torch.manual_seed(0)
m = InT.InT(dimensions=32, timesteps=T, kernel_size=7).cuda()
x, y = synth_pathtracker(B, T, HW, HW, n_distractors=14, device="cuda")
for _ in range(steps):
    m.zero_grad(set_to_none=True)
    out, _ = m(x)
    F.binary_cross_entropy_with_logits(out, y).backward()
torch.cuda.synchronize()
print("done", float(out.sum()))
