"""small fwd+bwd of every model family + the conv7 operator (run under compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathtracker_models_b200 import InT, ffhgru_hierarchy, synth_pathtracker
from pathtracker_models_b200.functional import conv7
torch.manual_seed(0)
for mode in (1, 0, 2):
    for ctor, kw in ((InT.InT, {}), (ffhgru_hierarchy.FFhGRU, {}), (InT.InT, {"no_inh": True})):
        m = ctor(32, 3, 7, **kw).cuda(); m.conv_mode = mode
        x, y = synth_pathtracker(2, 3, 32, 64)
        x.requires_grad_(True)
        out, _ = m(x); out.sum().backward()
    xx = torch.randn(2, 32, 32, 64, device="cuda"); w = torch.randn(32, 32, 7, 7, device="cuda") * 0.05
    conv7(xx, w, conv_mode=min(mode, 1), want_stats=True); conv7(xx, w, conv_mode=min(mode, 1), transpose=True)
torch.cuda.synchronize(); print("sanitize_small done")
