"""Fixture loading + the parity rule shared by the CPU and GPU suites."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# per tensor, max|a-b| / max|b| against the fp64 reference must be
# <= REL_TOL, or <= NOISE_MULT x the fp32 reference's own error vs fp64 for that tensor.
REL_TOL = 1e-4
NOISE_MULT = 4.0


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    B, T, H, W = [int(v) for v in z["shape"]]
    rs = np.random.RandomState(int(z["x_seed"]))
    x = rs.rand(B, 3, T, H, W)
    y = (rs.rand(B, 1) > 0.5).astype(np.float64)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    grads = {k[5:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("grad/")}
    noise = {k[7:]: float(z[k]) for k in z.files if k.startswith("f32err/")}
    extra = {k: torch.from_numpy(z[k]) for k in z.files if k.startswith("tm_")}
    return dict(x=torch.from_numpy(x), y=torch.from_numpy(y), sd=sd, grads=grads, noise=noise,
                logits=torch.from_numpy(z["logits"]), loss=float(z["loss"]), extra=extra,
                shape=(B, T, H, W))


def rel_err(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-300)


def check_parity(case, logits, grads, tol=REL_TOL, mult=NOISE_MULT):
    """returns list of (name, err, allowed) failures; logits use key '__logits__'."""
    bad = []
    e = rel_err(logits, case["logits"])
    allowed = max(tol, mult * case["noise"].get("__logits__", 0.0))
    if not e <= allowed:
        bad.append(("logits", e, allowed))
    for n, g_ref in case["grads"].items():
        if n not in grads or grads[n] is None:
            bad.append((n, float("nan"), 0.0))
            continue
        e = rel_err(grads[n], g_ref)
        allowed = max(tol, mult * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    return bad
