"""time the ConvLSTM forward loop (eval mode, no penalty): per-timestep cost = slope between two T values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pathtracker_models_b200 import convlstm
rs = np.random.RandomState(0)
gabor = torch.from_numpy(rs.normal(0, 0.3, (25, 1, 7, 7)).astype(np.float32))
x = torch.rand(64, 1, 32, 32, device="cuda")
tgt = torch.zeros(64, 32, 32, dtype=torch.long, device="cuda")
crit = torch.nn.CrossEntropyLoss()
res = {}
for T in (8, 24):
    m = convlstm.ConvLSTM(timesteps=T, filt_size=15, gabor=gabor).cuda().eval()
    with torch.no_grad():
        for _ in range(3): m(x, 0, 0, tgt, crit)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m(x, 0, 0, tgt, crit)
        e1.record(); torch.cuda.synchronize()
    res[T] = e0.elapsed_time(e1) / 5
print("fwd ms", res, "per-timestep us", (res[24] - res[8]) / 16 * 1e3)
