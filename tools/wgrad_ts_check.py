"""the sequence-summed 7x7 weight gradient (ptk_wgrad7 -> ptk_wgrad_ts.cu) alone: error vs torch fp64 + time per launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from pathtracker_models_b200.functional import wgrad7

def ref(inp, dy):
    w = torch.zeros(32, 32, 7, 7, dtype=torch.float64, device=inp.device, requires_grad=True)
    tot = 0
    for t in range(inp.shape[0]):
        tot = tot + (F.conv2d(inp[t].double(), w, padding=3) * dy[t].double()).sum()
    tot.backward()
    return w.grad

shapes = [(1, 1, 16, 32), (1, 2, 32, 32), (3, 3, 16, 64), (2, 2, 64, 64), (5, 7, 32, 32), (2, 64, 32, 32)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split(","))]
for (T, B, H, W) in shapes:
    g = torch.Generator(device="cuda").manual_seed(T * 100 + B)
    inp = torch.randn(T, B, 32, H, W, device="cuda", generator=g) * torch.rand(1, 1, 32, 1, 1, device="cuda", generator=g) * 2
    dy = torch.randn(T, B, 32, H, W, device="cuda", generator=g) * 0.1
    for mode in (1, 2):
        got = wgrad7(inp, dy, conv_mode=mode)
        torch.cuda.synchronize()
        r = ref(inp, dy)
        err = (got.double() - r).abs().max().item() / r.abs().max().item()
        # per-kernel-row error localises a layout mistake
        rows = [(got.double() - r)[:, :, ky].abs().max().item() / r.abs().max().item() for ky in range(7)]
        cols = [(got.double() - r)[:, :, :, kx].abs().max().item() / r.abs().max().item() for kx in range(7)]
        print(f"T={T} B={B} {H}x{W} mode={mode}: rel err {err:.2e}  by ky {['%.0e' % v for v in rows]} by kx {['%.0e' % v for v in cols]}", flush=True)
        assert torch.equal(got, wgrad7(inp, dy, conv_mode=mode)), "not deterministic"
T, B, H, W = 64, 64, 32, 32
inp = torch.randn(T, B, 32, H, W, device="cuda"); dy = torch.randn(T, B, 32, H, W, device="cuda")
for mode in (1, 2):
    for _ in range(2):
        wgrad7(inp, dy, conv_mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        wgrad7(inp, dy, conv_mode=mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cfg-2 sequence (T=64, B=64, 32x32) mode={mode}: {ms:.3f} ms per launch = {ms / 64 * 1e3:.1f} us per time step; "
          f"{2 * 64 * 64 * 1024 * 32 * 32 * 49 / ms / 1e9:.0f} TFLOP/s algorithmic")
