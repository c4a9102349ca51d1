"""The persistent forward time-loop kernels (one cooperative launch for all frames, replacing the Python loop of
models/InT.py:223-235) against the per-phase PDL chain and the reference fixtures.

PTK_FWD_LOOP=1 (csrc/ptk_int_loop.cu) runs the chain's own per-tile instruction sequences, so logits, testmode dumps, every gradient
and the eval-mode outputs must be BIT-identical to the chain.  PTK_FWD_LOOP=2 (csrc/ptk_int_loop2.cu: gate products on tcgen05,
thread = pixel) changes the arithmetic of the 1x1 gate products, so it is held to the same gate as every other path: the fp64
reference fixtures (1e-4 or 4x the fp32 reference's own error per tensor), plus agreement with the chain on the inference outputs.
The switch is read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, torch
sys.path.insert(0, %r)
from tests.golden_util import load_case, check_parity
from pathtracker_models_b200 import InT, ffhgru_hierarchy, _lib
res = {"bad": []}
for name, kind, kw, mode in [("int_tile", "int", {}, "bf16x3"), ("int_cfg1", "int", {}, "bf16x3"), ("int_64px", "int", {}, "bf16x3"),
                             ("int_tanh_small", "int", {"nl": torch.tanh}, "bf16x3"), ("hgru_tile", "hgru", {}, "bf16x3"),
                             ("int_small", "int", {}, "bf16x3"), ("hgru_small", "hgru", {}, "bf16x3"),
                             ("int_t64", "int", {}, "bf16x3"), ("hgru_t64", "hgru", {}, "bf16x3"), ("int_t128_64px", "int", {}, "bf16x3"),
                             ("int_tile", "int", {}, "bf16"), ("hgru_small", "hgru", {}, "bf16")]:
    case = load_case(name)
    m = (InT.InT(dimensions=32, timesteps=8, kernel_size=7, **kw) if kind == "int"
         else ffhgru_hierarchy.FFhGRU(dimensions=32, timesteps=8, kernel_size=7))
    m.load_state_dict(case["sd"]); m = m.cuda()
    m.conv_mode = _lib.PTK_CONV_BF16X3 if mode == "bf16x3" else _lib.PTK_CONV_BF16
    x = case["x"].float().cuda(); y = case["y"].float().cuda()
    outs = []
    for rep in range(2):                       # training step (checkpoints written by the forward feed the reverse loop)
        m.zero_grad(set_to_none=True)
        out, _ = m(x)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        outs.append(torch.cat([out.reshape(-1)] + [p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu())
    assert torch.equal(outs[0], outs[1]), name     # bit-reproducible
    res[name + "/" + mode + "/train"] = outs[0]
    if mode == "bf16x3":                       # the fp32-parity mode: the gate of tests/test_parity_gpu.py
        grads = {n: p.grad for n, p in m.named_parameters()}
        res["bad"] += [(name,) + tuple(b) for b in check_parity(case, out, grads)]
        if abs(loss.item() - case["loss"]) >= 1e-5:
            res["bad"].append((name, "loss", loss.item(), case["loss"]))
    with torch.no_grad():                      # inference: two-slot state ring instead of checkpoints
        m.eval()
        res[name + "/" + mode + "/eval"] = m(x)[0].cpu()
        o, states, gates = m(x, testmode=True)
        res[name + "/" + mode + "/tm"] = torch.cat([o.reshape(-1).cpu(), states.reshape(-1).cpu(), gates.reshape(-1).cpu()])
res["launches"] = torch.tensor(float(_lib.lib().ptk_launch_count()))
torch.save(res, sys.argv[1])
''' % ROOT


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    d = tmp_path_factory.mktemp("loop")
    res = {}
    for tag in ("0", "1", "2"):
        f = str(d / (tag + ".pt"))
        e = dict(os.environ); e["PTK_FWD_LOOP"] = tag
        r = subprocess.run([sys.executable, "-c", SCRIPT, f], env=e, capture_output=True, text=True, timeout=420)
        assert r.returncode == 0, r.stderr[-3000:]
        res[tag] = torch.load(f)
    return res


def test_chain_baseline_passes_the_reference_gate(runs):
    assert not runs["0"]["bad"], runs["0"]["bad"]


def test_persistent_forward_loop_is_bit_identical_to_the_chain(runs):
    assert runs["1"]["launches"] < runs["0"]["launches"]        # the loop kernel really ran
    for k, v in runs["0"].items():
        if k not in ("launches", "bad"):
            assert torch.equal(v, runs["1"][k]), k


def test_tcgen05_gate_loop_passes_the_reference_gate(runs):
    """second-generation loop kernel: logits, loss and every gradient against the fp64 reference fixtures (short cases, T = 64,
    T = 128 at 64x64, InT and FFhGRU); inference / testmode outputs within 1e-4 of the chain's."""
    assert runs["2"]["launches"] < runs["0"]["launches"]
    assert not runs["2"]["bad"], runs["2"]["bad"]
    for k, v in runs["0"].items():
        if k.endswith("/eval") or k.endswith("/tm"):
            w = runs["2"][k]
            tol = 1e-4 if "/bf16x3/" in k else 2e-2
            err = (v.double() - w.double()).abs().max().item() / max(v.double().abs().max().item(), 1e-30)
            assert err < tol, (k, err)
