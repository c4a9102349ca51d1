"""ConvLSTM tensor-core convolutions in isolation: one forward step and one backward step against the fp32 path.
Usage (GPU box): python tools/convk_check.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathtracker_models_b200 import _lib, convlstm
from tests.convlstm_util import load_convlstm_case

for name in ("convlstm_k7_small", "convlstm_k15_small"):
    case = load_convlstm_case(name)
    B, T, H, W = case["shape"]
    res = {}
    for mode in ("fp32", "bf16x3"):
        m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"])
        m.conv_mode = _lib.PTK_CONV_FP32 if mode == "fp32" else _lib.PTK_CONV_BF16X3
        m.load_state_dict(case["sd"])
        m = m.cuda().train()
        out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
        loss.backward()
        torch.cuda.synchronize()
        res[mode] = (out.detach(), pen.detach(), {n: p.grad.clone() for n, p in m.named_parameters()})
    a, b = res["fp32"], res["bf16x3"]
    rel = lambda x, y: ((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()
    print(name, "out", rel(b[0], a[0]), "pen", rel(b[1], a[1]))
    for n in a[2]:
        print("   ", n, rel(b[2][n], a[2][n]))
