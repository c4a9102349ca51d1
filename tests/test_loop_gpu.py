"""The persistent forward time-loop kernel (csrc/ptk_int_loop.cu: one cooperative launch for all frames, replacing the Python loop
of models/InT.py:223-235) against the per-phase PDL chain.  Both run the same per-tile instruction sequences, so logits, testmode
dumps, every checkpoint-dependent gradient and the eval-mode outputs must be BIT-identical; the chain itself is held to the
reference fixtures by test_parity_gpu.py.  The switch (PTK_FWD_LOOP) is read once per process, hence the subprocesses."""
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
from tests.golden_util import load_case
from pathtracker_models_b200 import InT, ffhgru_hierarchy, _lib
res = {}
for name, kind, kw, mode in [("int_tile", "int", {}, "bf16x3"), ("int_cfg1", "int", {}, "bf16x3"), ("int_64px", "int", {}, "bf16x3"),
                             ("int_tanh_small", "int", {"nl": torch.tanh}, "bf16x3"), ("hgru_tile", "hgru", {}, "bf16x3"),
                             ("int_t64", "int", {}, "bf16x3"), ("int_tile", "int", {}, "bf16"), ("hgru_small", "hgru", {}, "bf16")]:
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
        torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
        outs.append(torch.cat([out.reshape(-1)] + [p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu())
    assert torch.equal(outs[0], outs[1]), name
    res[name + "/" + mode + "/train"] = outs[0]
    with torch.no_grad():                      # inference: two-slot state ring instead of checkpoints
        m.eval()
        res[name + "/" + mode + "/eval"] = m(x)[0].cpu()
        o, states, gates = m(x, testmode=True)
        res[name + "/" + mode + "/tm"] = torch.cat([o.reshape(-1).cpu(), states.reshape(-1).cpu(), gates.reshape(-1).cpu()])
res["launches"] = torch.tensor(float(_lib.lib().ptk_launch_count()))
torch.save(res, sys.argv[1])
''' % ROOT


def test_persistent_forward_loop_is_bit_identical_to_the_chain(tmp_path):
    res = {}
    for tag, env in (("chain", {"PTK_FWD_LOOP": "0"}), ("loop", {"PTK_FWD_LOOP": "1"})):
        f = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", SCRIPT, f], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        res[tag] = torch.load(f)
    assert res["loop"]["launches"] < res["chain"]["launches"]        # the loop kernel really ran
    for k, v in res["chain"].items():
        if k != "launches":
            assert torch.equal(v, res["loop"][k]), k
