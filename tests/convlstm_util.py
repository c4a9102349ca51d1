"""ConvLSTM fixture loading shared by the CPU and GPU suites."""
import os

import numpy as np
import torch

from tests.golden_util import GOLDEN


def convlstm_shapes(filt):
    sh = {"conv0.weight": (25, 1, 7, 7), "conv0.bias": (25,)}
    for k in "ifco":
        sh[f"unit1.Wx{k}.weight"] = (25, 25, filt, filt)
        sh[f"unit1.Wx{k}.bias"] = (25,)
        sh[f"unit1.Wh{k}.weight"] = (25, 25, filt, filt)
    sh.update({"bn.weight": (25,), "bn.bias": (25,), "conv6.weight": (2, 25, 1, 1), "conv6.bias": (2,)})
    # state_dict order of the reference (models/convlstm.py:58-114)
    order = ["conv0.weight", "conv0.bias"]
    for k in "ifco":
        order += [f"unit1.Wx{k}.weight", f"unit1.Wx{k}.bias", f"unit1.Wh{k}.weight"]
    order += ["bn.weight", "bn.bias", "conv6.weight", "conv6.bias"]
    return {k: sh[k] for k in order}


def load_convlstm_case(name):
    from oracle.gen_golden import convlstm_weights
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    B, T, H, W = [int(v) for v in z["shape"]]
    filt = int(z["filt"])
    rs = np.random.RandomState(int(z["x_seed"]))
    x = rs.rand(B, 1, H, W)
    tgt = (rs.rand(B, H, W) > 0.5).astype(np.int64)
    if any(k.startswith("sd/") for k in z.files):
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    else:
        like = {k: torch.zeros(s) for k, s in convlstm_shapes(filt).items()}
        like["conv0.weight"] = torch.from_numpy(z["gabor"])
        sd = convlstm_weights(like, int(z["wseed"]))
    return dict(x=torch.from_numpy(x), target=torch.from_numpy(tgt), sd=sd, filt=filt, shape=(B, T, H, W),
                output=torch.from_numpy(z["output"]), loss=float(z["loss"]), penalty=torch.from_numpy(z["penalty"]),
                grads={k[5:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("grad/")},
                proj={k[5:]: z[k] for k in z.files if k.startswith("proj/")},
                gmax={k[5:]: float(z[k]) for k in z.files if k.startswith("gmax/")},
                noise={k[7:]: float(z[k]) for k in z.files if k.startswith("f32err/")},
                gabor=torch.from_numpy(z["gabor"]), proj_seed=int(z["x_seed"]) + 2)


def oracle_run(case):
    """fp64 oracle forward + autograd: (output, loss, penalty, grads keyed like the state_dict)."""
    from oracle import convlstm_oracle as CO
    p = CO.unpack(case["sd"])
    for v in p.values():
        v.requires_grad_(True)
    T = case["shape"][1]
    out, _ = CO.forward(case["x"], p, T)
    loss = torch.nn.functional.cross_entropy(out, case["target"])
    loss.backward()
    grads = {CO.STATE_KEYS[k]: v.grad for k, v in p.items()}
    with torch.enable_grad():
        pen = CO.jv_penalty(case["x"], {k: v.detach() for k, v in p.items()}, T)
    return out.detach(), loss.item(), pen.detach(), grads
