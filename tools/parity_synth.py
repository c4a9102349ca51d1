"""parity of the CUDA path on the BENCHMARK's own input distribution (synthetic PathTracker clips, default-init weights),
against the CPU port in fp64: logits, loss and every gradient, per conv mode.  Usage: parity_synth.py [B] [T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import int_oracle as O
from oracle.synth_oracle import synth
from pathtracker_models_b200 import InT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
m = InT.InT(dimensions=32, timesteps=T, kernel_size=7)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
xn, yn = synth(1234, 0, B, T, 32, 32, 14, 1)
res = {}
for dt in (torch.float64, torch.float32):
    p = O.unpack(sd, dtype=dt)
    for v in p.values():
        v.requires_grad_(True)
    lo, _ = O.forward(torch.from_numpy(xn).to(dt), p, "int", keep=False)
    loss = F.binary_cross_entropy_with_logits(lo, torch.from_numpy(yn).view(B, 1).to(dt))
    loss.backward()
    res[dt] = (lo.detach().double(), loss.item(), O.grad_to_state_keys({k: v.grad for k, v in p.items() if v.grad is not None}, sd))
l64, loss64, g64 = res[torch.float64]
rel = lambda a, b: (a.double().cpu() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-300)
print("fp64 logits", l64.flatten()[:4].tolist(), "loss", loss64)
print("cpu fp32 port : logits rel %.2e  loss abs %.2e  worst grad %.2e" % (
    rel(res[torch.float32][0], l64), abs(res[torch.float32][1] - loss64), max(rel(res[torch.float32][2][n], g64[n]) for n in g64 if float(g64[n].abs().max()) > 0)))
m = m.cuda()
x, y = torch.from_numpy(xn).cuda(), torch.from_numpy(yn).view(B, 1).cuda()
for mode, name in ((0, "fp32"), (1, "bf16x3")):
    m.conv_mode = mode
    m.zero_grad(set_to_none=True)
    out, _ = m(x)
    loss = F.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    errs = sorted(((rel(p.grad, g64[n]), n, float(g64[n].abs().max())) for n, p in m.named_parameters() if p.grad is not None and n in g64 and float(g64[n].abs().max()) > 0), reverse=True)
    print("%-7s: logits rel %.2e  loss abs %.2e   worst grads: %s" % (name, rel(out, l64), abs(loss.item() - loss64),
          ", ".join("%s %.1e (max|g| %.1e)" % (n, e, mx) for e, n, mx in errs[:5])))
