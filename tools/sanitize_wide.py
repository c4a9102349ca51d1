"""small fwd+bwd of the ConvLSTM (filt 7 and 15, bptt and rbp) and kys.GRU in the tensor-core mode (run under
compute-sanitizer: the wide tcgen05 kernels of ptk_convk_tc.cu / ptk_wgradk_tc.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pathtracker_models_b200 import convlstm, kys
torch.manual_seed(0)
gabor = torch.from_numpy(np.random.RandomState(0).normal(0, 0.3, (25, 1, 7, 7)).astype(np.float32))
crit = torch.nn.CrossEntropyLoss()
for filt, H, W, gm in ((7, 32, 32, 'bptt'), (15, 16, 64, 'bptt'), (7, 32, 32, 'rbp')):
    m = convlstm.ConvLSTM(timesteps=3, filt_size=filt, gabor=gabor, grad_method=gm, num_iter=3).cuda().train()
    x = torch.rand(3, 1, H, W, device="cuda"); tgt = (torch.rand(3, H, W, device="cuda") > 0.5).long()
    out, pen, loss = m(x, 0, 0, tgt, crit); loss.backward()
g = kys.GRU(dimensions=64, timesteps=2, kernel_size=7).cuda()
x = torch.rand(3, 3, 2, 16, 64, device="cuda", requires_grad=True)
out, _ = g(x); out.sum().backward()
torch.cuda.synchronize(); print("sanitize_wide done")
