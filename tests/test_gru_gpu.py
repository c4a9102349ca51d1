"""GPU parity of the ConvGRU path (models/kys.py, `--model gru`) through the C ABI, against fixtures produced by the
reference and the fp64 oracle."""
import numpy as np
import pytest
import torch

from tests.golden_util import NOISE_MULT, REL_TOL, rel_err
from tests.test_gru_oracle import load_gru_case, mirror_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", ["gru_small", "gru_b3"])
def test_gru_logits_and_all_grads_vs_reference(name, conv_mode):
    """both precisions: 32x32-channel blocks on the tcgen05 kernels (default) and the generic strict-fp32 kernels"""
    from oracle import gru_oracle as GO
    from pathtracker_models_b200 import _lib
    case = load_gru_case(name)
    m = mirror_model(case).cuda()
    m.conv_mode = _lib.PTK_CONV_BF16X3 if conv_mode == "bf16x3" else _lib.PTK_CONV_FP32
    x, y = case["x"].float().cuda(), case["y"].float().cuda()
    out, pen = m(x)
    assert out.shape == (case["shape"][0], 1) and pen.shape == (1,)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    assert rel_err(out, case["logits"]) < REL_TOL and abs(loss.item() - case["loss"]) < 1e-5
    # fp64 oracle gradients (the oracle itself is pinned to the reference on the CPU suite)
    p = {k: v.detach().double().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    torch.nn.functional.binary_cross_entropy_with_logits(GO.forward(case["x"], p), case["y"]).backward()
    bad = []
    for n, prm in m.named_parameters():
        if p[n].grad is None:
            assert prm.grad is None, n          # the unused BatchNorm3d (models/kys.py:80)
            continue
        e = rel_err(prm.grad, p[n].grad)
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)
    for n, g in case["grads"].items():          # and directly against the reference's own small tensors
        assert rel_err(dict(m.named_parameters())[n].grad, g) < max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0)), n


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
def test_gru_long_horizon_t64_vs_reference(conv_mode):
    """T = 64 (the horizon `bench.py --model gru` is quoted on), B = 2.  Fixture: the reference in fp64 (logits, loss, the
    small gradients in full, the three 7x7 banks through seeded projections).  The per-tensor check of the big banks uses the
    fp64 oracle evaluated with torch on the GPU, after that evaluation has itself been held to the fixture."""
    from oracle import gru_oracle as GO
    from pathtracker_models_b200 import _lib
    case = load_gru_case("gru_t64")
    m = mirror_model(case).cuda()
    m.conv_mode = _lib.PTK_CONV_BF16X3 if conv_mode == "bf16x3" else _lib.PTK_CONV_FP32
    x, y = case["x"].float().cuda(), case["y"].float().cuda()
    out, _ = m(x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    assert rel_err(out, case["logits"]) < REL_TOL and abs(loss.item() - case["loss"]) < 1e-5
    p = {k: v.detach().double().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    lo = GO.forward(case["x"].cuda(), p)
    torch.nn.functional.binary_cross_entropy_with_logits(lo, case["y"].cuda()).backward()
    assert rel_err(lo, case["logits"]) < 1e-9                       # the checker against the reference's own output
    prs = np.random.RandomState(case["seed"] + 2)
    for n, prm in p.items():
        if prm.grad is None:
            continue
        if n in case["grads"]:
            assert rel_err(prm.grad, case["grads"][n]) < 1e-8, n
        else:
            g = prm.grad.cpu().numpy().ravel()
            v = prs.normal(0, 1, (4, g.size))
            assert np.allclose(v @ g, case["proj"][n], rtol=1e-7, atol=1e-11 * max(1.0, case["gmax"][n])), n
    bad = []
    for n, prm in m.named_parameters():
        if p[n].grad is None:
            continue
        e = rel_err(prm.grad, p[n].grad)
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)


def test_gru_testmode_input_gradient_and_determinism():
    case = load_gru_case("gru_small")
    m = mirror_model(case).cuda()
    x = case["x"].float().cuda().requires_grad_(True)
    out, states, gates = m(x, testmode=True)
    B, T, H, W = case["shape"]
    assert states.shape == (B, T, 1, H, W) and gates.shape == (B, T, 64, H, W)
    assert rel_err(states, torch.from_numpy(case["z"]["tm_states"])) < REL_TOL
    assert abs(gates.double().sum().item() - float(case["z"]["tm_gates_sum"])) < 1e-5 * abs(float(case["z"]["tm_gates_sum"]))
    assert rel_err(gates[:, -1, :4], torch.from_numpy(case["z"]["tm_gates_last"])) < REL_TOL
    out.sum().backward()
    g1 = x.grad.clone()
    # fp64 oracle input gradient
    from oracle import gru_oracle as GO
    p = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    x64 = case["x"].clone().requires_grad_(True)
    GO.forward(x64, p).sum().backward()
    p32 = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}     # the golden gate: 1e-4 or 4 x the fp32 oracle's own error
    x32 = case["x"].float().clone().requires_grad_(True)
    GO.forward(x32, p32).sum().backward()
    assert rel_err(g1, x64.grad) <= max(REL_TOL, NOISE_MULT * rel_err(x32.grad, x64.grad))
    x2 = case["x"].float().cuda().requires_grad_(True)
    out2, _ = m(x2)
    out2.sum().backward()
    assert torch.equal(out, out2) and torch.equal(g1, x2.grad)      # bit-exact reruns


def test_gru_errors_are_loud():
    from pathtracker_models_b200 import _lib, kys
    with pytest.raises(NotImplementedError):
        kys.GRU(dimensions=32, timesteps=4, kernel_size=7)
    m = kys.GRU(dimensions=64, timesteps=4, kernel_size=7)
    with pytest.raises(_lib.PtkError):
        m(torch.rand(1, 3, 2, 32, 32))                    # CPU tensor: no fallback
    with pytest.raises(_lib.PtkError):
        m.cuda()(torch.rand(1, 3, 2, 20, 32, device="cuda"))   # H % 16 != 0


def test_data_parallel_hooks_on_gru_and_convlstm_world1(tmp_path):
    """attach_data_parallel on the other families: with a 1-rank group the averaged gradients must equal the plain ones"""
    import torch.distributed as dist
    from pathtracker_models_b200.parallel import attach_data_parallel, broadcast_parameters
    case = load_gru_case("gru_small")
    m = mirror_model(case).cuda()
    x, y = case["x"].float().cuda(), case["y"].float().cuda()
    out, _ = m(x)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="file://" + str(tmp_path / "pg"), rank=0, world_size=1)
    try:
        broadcast_parameters(m)
        attach_data_parallel(m)
        out2, _ = m(x)
        torch.nn.functional.binary_cross_entropy_with_logits(out2, y).backward()
        for n, p in m.named_parameters():
            if p.grad is not None:
                assert torch.equal(p.grad, ref[n]), n
    finally:
        if created:
            dist.destroy_process_group()
