"""CPU: pin the ConvGRU oracle (and the mirror module's init) against fixtures produced by the reference (models/kys.py)."""
import os

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN, rel_err


def load_gru_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    B, T, H, W = [int(v) for v in z["shape"]]
    seed = int(z["x_seed"])
    rs = np.random.RandomState(seed)
    x = rs.rand(B, 3, T, H, W)
    y = (rs.rand(B, 1) > 0.5).astype(np.float64)
    return dict(z=z, seed=seed, shape=(B, T, H, W), x=torch.from_numpy(x), y=torch.from_numpy(y),
                logits=torch.from_numpy(z["logits"]), loss=float(z["loss"]),
                grads={k[5:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("grad/")},
                proj={k[5:]: z[k] for k in z.files if k.startswith("proj/")},
                gmax={k[5:]: float(z[k]) for k in z.files if k.startswith("gmax/")},
                noise={k[7:]: float(z[k]) for k in z.files if k.startswith("f32err/")})


def mirror_model(case):
    """the mirror module under the fixture's seed: its default init must equal the reference's bit for bit"""
    from pathtracker_models_b200 import kys
    torch.manual_seed(case["seed"])
    return kys.GRU(dimensions=64, timesteps=case["shape"][1], kernel_size=7)


@pytest.mark.parametrize("name", ["gru_small", "gru_b3"])
def test_gru_init_is_bit_identical_and_oracle_matches_reference(name):
    from oracle import gru_oracle as GO
    case = load_gru_case(name)
    m = mirror_model(case)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in case["z"]["init_keys"]]
    assert np.array_equal(np.array([v.double().sum().item() for v in sd.values()]), case["z"]["init_sums"])
    p = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    out = GO.forward(case["x"], p)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, case["y"])
    loss.backward()
    assert rel_err(out, case["logits"]) < 1e-10 and abs(loss.item() - case["loss"]) < 1e-10
    for n, g in case["grads"].items():
        assert rel_err(p[n].grad, g) < 1e-9, n
    prs = np.random.RandomState(case["seed"] + 2)
    for n, prm in p.items():       # same draw order as the generator: only the big tensors consume random numbers
        if prm.grad is None or n in case["grads"]:
            continue
        g = prm.grad.numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert np.allclose(v @ g, case["proj"][n], rtol=1e-8, atol=1e-12 * max(1.0, case["gmax"][n])), n
    with torch.no_grad():
        _, states, gates = GO.forward(case["x"], {k: v.detach() for k, v in p.items()}, testmode=True)
    assert rel_err(states, torch.from_numpy(case["z"]["tm_states"])) < 1e-6
    assert abs(gates.sum().item() - float(case["z"]["tm_gates_sum"])) < 1e-6 * abs(float(case["z"]["tm_gates_sum"]))
