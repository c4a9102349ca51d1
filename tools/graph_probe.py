"""eager vs CUDA-graph replay of one fwd+bwd step (launch-gap probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathtracker_models_b200 import InT, synth_pathtracker
torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = InT.InT(32, T, 7).cuda()
x, y = synth_pathtracker(64, T, 32, 32)
def step():
    for p in m.parameters(): p.grad = None
    out, _ = m(x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    return loss
def timeit(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for _ in range(3): step()
print("eager  %.2f ms/step" % timeit(step))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = step()
print("graph  %.2f ms/step" % timeit(g.replay), "loss", float(loss))
