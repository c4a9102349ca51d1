"""GPU parity of the ConvLSTM path (through the C ABI) against fixtures from the reference and the fp64 oracle."""
import pytest
import torch

from tests.convlstm_util import convlstm_shapes, load_convlstm_case, oracle_run
from tests.golden_util import NOISE_MULT, REL_TOL, rel_err

pytestmark = pytest.mark.gpu


def _mode(conv_mode):
    from pathtracker_models_b200 import _lib
    return _lib.PTK_CONV_BF16X3 if conv_mode == "bf16x3" else _lib.PTK_CONV_FP32


def _run(case, conv_mode="bf16x3"):
    from pathtracker_models_b200 import convlstm
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"])
    m.conv_mode = _mode(conv_mode)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    return m, out, pen, loss


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", ["convlstm_k7_small", "convlstm_k15_small", "convlstm_k15_b3"])
def test_convlstm_output_penalty_and_grads(name, conv_mode):
    """both precisions of the per-step convolutions: tcgen05 split-bf16 (default) and strict fp32 on CUDA cores"""
    case = load_convlstm_case(name)
    m, out, pen, loss = _run(case, conv_mode)
    B, T, H, W = case["shape"]
    assert out.shape == (B, 2, H, W) and pen.shape == (B, 25, H, W)
    assert rel_err(out, case["output"]) < REL_TOL
    assert abs(loss.item() - case["loss"]) < 1e-5
    assert rel_err(pen, case["penalty"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__penalty__", 0))
    _, _, _, g64 = oracle_run(case)                 # fp64 oracle (itself pinned to the reference on the CPU suite)
    bad = []
    for n, p in m.named_parameters():
        e = rel_err(p.grad, g64[n])
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
def test_convlstm_long_horizon_k15_t64_vs_reference(conv_mode):
    """BASELINE configs[4] at its own horizon: filt 15, timesteps = 64 (models/convlstm.py:138), B = 2, train mode (penalty
    branch).  Fixture = the reference in fp64 (output, loss, penalty, every gradient through seeded projections); the
    per-tensor gradient check uses the fp64 oracle evaluated with torch on the GPU, itself held to those projections."""
    import numpy as np
    from oracle import convlstm_oracle as CO
    case = load_convlstm_case("convlstm_k15_t64")
    m, out, pen, loss = _run(case, conv_mode)
    B, T, H, W = case["shape"]
    assert T == 64 and out.shape == (B, 2, H, W)
    assert rel_err(out, case["output"]) < REL_TOL
    assert abs(loss.item() - case["loss"]) < 1e-5
    assert rel_err(pen, case["penalty"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__penalty__", 0))
    p = {k: v.cuda().requires_grad_(True) for k, v in CO.unpack(case["sd"]).items()}
    o64, _ = CO.forward(case["x"].cuda(), p, T)
    torch.nn.functional.cross_entropy(o64, case["target"].cuda()).backward()
    g64 = {CO.STATE_KEYS[k]: v.grad for k, v in p.items()}
    assert rel_err(o64, case["output"]) < 1e-9
    prs = np.random.RandomState(case["proj_seed"])
    for n in case["proj"]:                                  # the checker against the reference's own gradients
        g = g64[n].cpu().numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert np.allclose(v @ g, case["proj"][n], rtol=1e-7, atol=1e-11 * max(1.0, case["gmax"][n])), n
    bad = []
    for n, prm in m.named_parameters():
        e = rel_err(prm.grad, g64[n])
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)


def test_convlstm_eval_mode_and_determinism():
    from pathtracker_models_b200 import convlstm
    case = load_convlstm_case("convlstm_k7_small")
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=7, gabor=case["gabor"])
    m.load_state_dict(case["sd"])
    m = m.cuda().eval()
    x, tgt = case["x"].float().cuda(), case["target"].cuda()
    with torch.no_grad():
        o1, pen, l1 = m(x, 0, 0, tgt, torch.nn.CrossEntropyLoss())
        o2, states, l2 = m(x, 0, 0, tgt, torch.nn.CrossEntropyLoss(), testmode=True)
    assert pen.shape == (1,) and states == [] and torch.equal(o1, o2)
    assert rel_err(o1, case["output"]) < REL_TOL


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", ["convlstm_k7_rbp", "convlstm_k7_rbp_trunc"])
def test_convlstm_rbp_grads_vs_reference(name, conv_mode):
    """recurrent back-propagation: CUDA Neumann series vs the reference's own outputs."""
    import os
    import numpy as np
    from oracle import convlstm_oracle as CO
    from pathtracker_models_b200 import convlstm
    from tests.golden_util import GOLDEN
    case = load_convlstm_case(name)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    num_iter = int(z["num_iter"])
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"], grad_method='rbp', num_iter=num_iter)
    m.conv_mode = _mode(conv_mode)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    assert rel_err(out, case["output"]) < REL_TOL and abs(loss.item() - case["loss"]) < 1e-5
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in case["sd"].items()}
    _, _, g64, terms = CO.rbp_grads(case["x"], p64, T, case["target"], num_iter)
    assert int(m.last_rbp_terms.item()) == terms
    bad = []
    for n, p in m.named_parameters():
        e = rel_err(p.grad, case["grads"][n])
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)
    # testmode of the rbp branch returns the T+1 hidden states (models/convlstm.py:127-135)
    m.eval()
    with torch.no_grad():
        o2, states, _ = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss(), testmode=True)
    assert len(states) == T + 1 and states[0].abs().max().item() == 0.0 and states[-1].shape == (B, 25, H, W)


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("filt,H,W", [(15, 32, 64), (7, 64, 64)])
def test_convlstm_wide_frames_vs_fp64_oracle(filt, H, W, conv_mode):
    """frames wider / taller than one 32-column tile: the halo columns and rows of the convolution tiles carry real
    neighbours (the 32x32 fixtures only ever see zero padding there).  Checked against the fp64 oracle, which the CPU
    suite pins to the reference's own outputs."""
    import numpy as np
    from oracle.gen_golden import convlstm_weights
    from pathtracker_models_b200 import convlstm
    base = load_convlstm_case("convlstm_k7_small")
    B, T = 2, 3
    rs = np.random.RandomState(77)
    like = {k: torch.zeros(s) for k, s in convlstm_shapes(filt).items()}
    like["conv0.weight"] = base["gabor"].reshape(25, 1, 7, 7).clone()
    case = dict(x=torch.from_numpy(rs.rand(B, 1, H, W)), target=torch.from_numpy((rs.rand(B, H, W) > 0.5).astype(np.int64)),
                sd=convlstm_weights(like, 5), filt=filt, shape=(B, T, H, W))
    m = convlstm.ConvLSTM(timesteps=T, filt_size=filt, gabor=base["gabor"])
    m.conv_mode = _mode(conv_mode)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    o64, l64, p64, g64 = oracle_run(case)
    assert rel_err(out, o64) < REL_TOL and abs(loss.item() - l64) < 1e-5
    assert rel_err(pen, p64) < 4 * REL_TOL          # the penalty is a clamp of a Jacobian product: compare loosely
    bad = [(n, rel_err(p.grad, g64[n])) for n, p in m.named_parameters() if not rel_err(p.grad, g64[n]) <= 2 * REL_TOL]
    assert not bad, bad


@pytest.mark.parametrize("K,B,H,W", [(15, 64, 32, 32), (7, 64, 32, 32), (15, 3, 16, 64), (7, 1, 64, 32), (15, 1, 8, 32)])
def test_wide_operators_alone_vs_torch_fp64(K, B, H, W):
    """the three tcgen05 operators of the time loops on their own (through ptk_lstm_wide_op) against torch fp64
    convolutions, at BASELINE's full batch and at ragged shapes (B = 1: fewer tiles than weight-gradient CTAs)."""
    import torch.nn.functional as F
    from pathtracker_models_b200.functional import lstm_wide_op
    g = torch.Generator().manual_seed(K * 1000 + B)
    w = torch.randn(4, 25, 25, K, K, generator=g, dtype=torch.float64) / (25 * K * K) ** 0.5
    h = (torch.rand(B, 25, H, W, generator=g, dtype=torch.float64) * 2 - 1)
    dy = torch.randn(B, 100, H, W, generator=g, dtype=torch.float64)
    add = torch.randn(B, 100, H, W, generator=g, dtype=torch.float64)
    w32, h32, dy32, add32 = (t.float() for t in (w, h, dy, add))        # the fp64 reference sees the fp32-rounded inputs
    w64, h64, dy64, add64 = (t.double().requires_grad_(t is not add32) for t in (w32, h32, dy32, add32))
    ws = w64.reshape(100, 25, K, K)
    ref = F.conv2d(h64, ws, padding=K // 2) + add64
    got = lstm_wide_op("conv", h32.cuda(), w32.cuda(), add32.cuda())
    tol = 5e-5      # the 3-pass bf16 split carries 16-17 mantissa bits per operand; measured 1e-5 ... 3e-5 of max here
    assert rel_err(got, ref.detach()) < tol
    (F.conv2d(h64, ws, padding=K // 2) * dy64.detach()).sum().backward()
    assert rel_err(lstm_wide_op("transpose", dy32.cuda(), w32.cuda()), h64.grad) < tol
    assert rel_err(lstm_wide_op("wgrad", h32.cuda(), w32.cuda(), dy32.cuda()), w64.grad) < tol


@pytest.mark.parametrize("grad_method", ["bptt", "rbp"])
def test_convlstm_input_gradient_vs_fp64_oracle(grad_method):
    """the reference is differentiable w.r.t. the image through autograd (x feeds conv0, models/convlstm.py:118):
    x.grad = conv0^T (d loss / d u), ptk_convlstm_input_grad.  BPTT: against autograd through the fp64 oracle;
    rbp: x.grad must equal conv_transpose2d of the conv0.weight-gradient's own d(loss)/d(u), checked through the identity
    <x.grad, x> = <conv0.weight.grad, conv0.weight> (both equal <du, conv0(x) - b>)."""
    from oracle import convlstm_oracle as CO
    from pathtracker_models_b200 import convlstm
    case = load_convlstm_case("convlstm_k7_small")
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"], grad_method=grad_method, num_iter=4)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    x = case["x"].float().cuda().requires_grad_(True)
    out, pen, loss = m(x, 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    assert x.grad is not None and x.grad.shape == x.shape and torch.isfinite(x.grad).all()
    if grad_method == "bptt":
        p = {k: v.cuda() for k, v in CO.unpack(case["sd"]).items()}
        x64 = case["x"].double().cuda().requires_grad_(True)
        o64, _ = CO.forward(x64, p, T)
        torch.nn.functional.cross_entropy(o64, case["target"].cuda()).backward()
        # the usual gate: 1e-4 of max, or 4x the error of the same computation in fp32 (oracle port on the CPU, no TF32)
        p32 = CO.unpack(case["sd"], dtype=torch.float32)
        x32 = case["x"].float().requires_grad_(True)
        o32, _ = CO.forward(x32, p32, T)
        torch.nn.functional.cross_entropy(o32, case["target"]).backward()
        allowed = max(REL_TOL, NOISE_MULT * rel_err(x32.grad, x64.grad))
        assert rel_err(x.grad, x64.grad) <= allowed, (rel_err(x.grad, x64.grad), allowed)
    lhs = float((x.grad.double() * x.detach().double()).sum())
    rhs = float((m.conv0.weight.grad.double() * m.conv0.weight.detach().double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1e-12), (lhs, rhs)


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", ["convlstm_k7_jp", "convlstm_k7_jp_t2", "convlstm_k15_jp", "convlstm_k7_rbp_jp", "convlstm_k15_jp_t64"])
def test_convlstm_jacobian_penalty_in_graph_vs_reference(name, conv_mode):
    """jacobian_penalty=True (create_graph=True, models/convlstm.py:157-160): the training loop back-propagates
    loss + 10 * penalty.mean() (mainclean.py:192-196).  The CUDA path takes the penalty's double back-propagation in closed
    form (ptk_convlstm_backward_penalty); fixtures = the reference's own gradients of that sum, both gradient methods."""
    import os
    import numpy as np
    from oracle import convlstm_oracle as CO
    from pathtracker_models_b200 import convlstm
    from tests.golden_util import GOLDEN
    case = load_convlstm_case(name)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    method = "rbp" if "rbp" in name else "bptt"
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"], grad_method=method,
                          num_iter=int(z["num_iter"]), jacobian_penalty=True)
    m.conv_mode = _mode(conv_mode)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    assert pen.requires_grad
    (loss + pen.mean() * 1e1).backward()
    # the gate of SURVEY 8c for the output too: at timesteps = 64 the fp32 reference itself is 8.5e-5 off its fp64 run on this fixture
    assert rel_err(out, case["output"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__output__", 0.0))
    assert abs(loss.item() - case["loss"]) < 1e-5
    assert rel_err(pen, case["penalty"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__penalty__", 0))
    # checker: the fp64 double back-propagation oracle on the GPU, itself pinned to the reference on the CPU suite
    p64 = {k: v.double().cuda().requires_grad_(True) for k, v in case["sd"].items()}
    _, _, _, g64 = CO.jp_total_grads(case["x"].cuda(), p64, T, case["target"].cuda(), grad_method=method, num_iter=int(z["num_iter"]))
    for n, g in case["grads"].items():
        assert rel_err(g64[n], g) < (2e-7 if g.numel() > 4096 else 1e-9), n
    prs = np.random.RandomState(case["proj_seed"])
    for n in case["proj"]:                                  # compact fixtures: the reference's gradients through seeded projections
        g = g64[n].detach().cpu().numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert np.allclose(v @ g, case["proj"][n], rtol=1e-7, atol=1e-11 * max(1.0, case["gmax"][n])), n
    bad = []
    for n, prm in m.named_parameters():
        e = rel_err(prm.grad, g64[n])
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)
    g_total = {n: prm.grad.clone() for n, prm in m.named_parameters()}
    # linearity: the penalty alone (the output's gradient arrives as None) + the loss alone = the sum
    m.zero_grad()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    (pen.mean() * 1e1).backward()
    g_pen = {n: prm.grad.clone() for n, prm in m.named_parameters()}
    assert max(float(g.abs().max()) for n, g in g_pen.items() if "unit1.Wh" in n) > 0
    assert all(float(g_pen[n].abs().max()) == 0.0 for n in ("bn.weight", "bn.bias", "conv6.weight", "conv6.bias"))   # the head is not in the penalty
    # a zero penalty gradient through the penalty entry point (d pen = 0 tensor: the seeds vanish) ...
    m.zero_grad()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    (loss + 0.0 * pen.sum()).backward()
    for n, prm in m.named_parameters():
        scale = float(g_total[n].abs().max())
        assert float((g_total[n] - prm.grad - g_pen[n]).abs().max()) <= REL_TOL * scale + 1e-12, n    # fp32 sums in a different grouping
    # ... equals the ordinary backward of a module without the flag
    m2 = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"], grad_method=method, num_iter=int(z["num_iter"]))
    m2.conv_mode = _mode(conv_mode)
    m2.load_state_dict(case["sd"])
    m2 = m2.cuda().train()
    _, pen2, loss2 = m2(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    assert not pen2.requires_grad and torch.equal(pen2, pen.detach())
    loss2.backward()
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert rel_err(a.grad, b.grad) < 1e-6, n     # d pen = 0: the seeds vanish, the reverse loop is the ordinary one


@pytest.mark.parametrize("name", ["convlstm_k7_rbp", "convlstm_k7_rbp_trunc"])
def test_convlstm_rbp_penalty_value_vs_reference(name):
    """grad_method='rbp': h_{T-1}, c_{T-1} are detached leaves (models/convlstm.py:130-131), so the second product is f_T alone"""
    import os
    import numpy as np
    from pathtracker_models_b200 import convlstm
    from tests.golden_util import GOLDEN
    case = load_convlstm_case(name)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"], grad_method='rbp', num_iter=int(z["num_iter"]))
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    _, pen, _ = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    assert rel_err(pen, case["penalty"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__penalty__", 0))


@pytest.mark.parametrize("method", ["bptt", "rbp"])
@pytest.mark.parametrize("filt,B,T,H,W", [(15, 3, 5, 32, 64), (7, 1, 8, 64, 32)])
def test_convlstm_jacobian_penalty_wide_frames_vs_fp64_oracle(filt, B, T, H, W, method):
    """the in-graph penalty on frames with real halo neighbours, an odd batch and a longer loop (the second-order seeds sit
    at steps T and T-1 only, whatever T is), with a penalty weight other than the training loop's 10 and the input gradient
    (saliency through the penalty).  Checker: fp64 double back-propagation (oracle, pinned to the reference on the CPU suite)."""
    import numpy as np
    from oracle import convlstm_oracle as CO
    from oracle.gen_golden import convlstm_weights
    from pathtracker_models_b200 import convlstm
    base = load_convlstm_case("convlstm_k7_small")
    rs = np.random.RandomState(78)
    like = {k: torch.zeros(s) for k, s in convlstm_shapes(filt).items()}
    like["conv0.weight"] = base["gabor"].reshape(25, 1, 7, 7).clone()
    sd = convlstm_weights(like, 6)
    x = torch.from_numpy(rs.rand(B, 1, H, W))
    tgt = torch.from_numpy((rs.rand(B, H, W) > 0.5).astype(np.int64))
    m = convlstm.ConvLSTM(timesteps=T, filt_size=filt, gabor=base["gabor"], grad_method=method, num_iter=6, jacobian_penalty=True)
    m.load_state_dict(sd)
    m = m.cuda().train()
    xg = x.float().cuda().requires_grad_(True)
    out, pen, loss = m(xg, 0, 0, tgt.cuda(), torch.nn.CrossEntropyLoss())
    (loss + 3.0 * pen.mean()).backward()
    p64 = {k: v.double().cuda().requires_grad_(True) for k, v in sd.items()}
    x64 = x.cuda().requires_grad_(True)
    o64, l64, pen64, g64 = CO.jp_total_grads(x64, p64, T, tgt.cuda(), grad_method=method, num_iter=6, weight=3.0)
    assert rel_err(out, o64) < REL_TOL and abs(loss.item() - l64) < 1e-5
    assert rel_err(pen, pen64) < 4 * REL_TOL
    assert float(pen64.max()) > 0
    # the same rule as the fixture cases: 1e-4, or 4 x the fp32 error of the reference algorithm itself on this input
    # (conv0.bias is a heavily cancelling sum: the fp32 oracle is 1e-4 off there on the filt-15 cases)
    p32 = {k: v.float().cuda().requires_grad_(True) for k, v in sd.items()}
    _, _, _, g32 = CO.jp_total_grads(x.float().cuda(), p32, T, tgt.cuda(), grad_method=method, num_iter=6, weight=3.0)
    bad = []
    for n, prm in m.named_parameters():
        e, allowed = rel_err(prm.grad, g64[n]), max(2 * REL_TOL, NOISE_MULT * rel_err(g32[n], g64[n]))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, bad
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and float(xg.grad.abs().max()) > 0
    if method == "bptt":
        assert rel_err(xg.grad, g64["__x__"]) <= 2 * REL_TOL


LSTM_BF16_OUT_TOL = 2e-3      # measured 1.7e-4 .. 3.5e-4
LSTM_BF16_GRAD_TOL = 3e-2     # measured worst 1.4e-2 (conv0.bias, filt 7), 6.8e-3 / 4.3e-3 (filt 15)


@pytest.mark.parametrize("name", ["convlstm_k7_small", "convlstm_k15_small", "convlstm_k15_b3"])
def test_convlstm_bf16_single_pass_mode_error_is_bounded(name):
    """the bf16 variant (conv_mode = PTK_CONV_BF16: one bf16 product per k-step in the wide tcgen05 kernels) is NOT the parity
    path: operands carry 8 mantissa bits.  Its error is bounded and documented here; it is opt-in."""
    from pathtracker_models_b200 import _lib, convlstm
    case = load_convlstm_case(name)
    B, T, H, W = case["shape"]
    m = convlstm.ConvLSTM(timesteps=T, filt_size=case["filt"], gabor=case["gabor"])
    m.conv_mode = _lib.PTK_CONV_BF16
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    _, _, _, g64 = oracle_run(case)
    eo = rel_err(out, case["output"])
    worst = max((rel_err(p.grad, g64[n]), n) for n, p in m.named_parameters())
    print(name, "bf16 single-pass: output", eo, "worst gradient", worst)
    assert eo < LSTM_BF16_OUT_TOL
    assert worst[0] < LSTM_BF16_GRAD_TOL, worst
    assert worst[0] > 1e-5          # the mode really ran reduced precision


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
def test_convlstm_rbp_long_horizon_k15_t64_vs_reference(conv_mode):
    """recurrent back-propagation at configs[4]'s horizon (filt 15, timesteps = 64, B = 2) on the constant-memory checkpoint ring:
    output / loss / penalty against the reference's own fixture, every gradient against the fp64 oracle (held to the fixture's
    projections on the CPU suite), the accepted Neumann terms against the oracle's count."""
    import os
    import numpy as np
    from oracle import convlstm_oracle as CO
    from pathtracker_models_b200 import convlstm
    from tests.golden_util import GOLDEN
    case = load_convlstm_case("convlstm_k15_rbp_t64")
    z = np.load(os.path.join(GOLDEN, "convlstm_k15_rbp_t64.npz"))
    num_iter = int(z["num_iter"])
    B, T, H, W = case["shape"]
    assert T == 64
    m = convlstm.ConvLSTM(timesteps=T, filt_size=15, gabor=case["gabor"], grad_method='rbp', num_iter=num_iter)
    m.conv_mode = _mode(conv_mode)
    m.load_state_dict(case["sd"])
    m = m.cuda().train()
    out, pen, loss = m(case["x"].float().cuda(), 0, 0, case["target"].cuda(), torch.nn.CrossEntropyLoss())
    loss.backward()
    assert rel_err(out, case["output"]) < REL_TOL and abs(loss.item() - case["loss"]) < 1e-5
    assert rel_err(pen, case["penalty"]) < max(REL_TOL, NOISE_MULT * case["noise"].get("__penalty__", 0))
    p64 = {k: v.double().cuda().requires_grad_(True) for k, v in case["sd"].items()}
    _, _, g64, terms = CO.rbp_grads(case["x"].cuda(), p64, T, case["target"].cuda(), num_iter)
    assert int(m.last_rbp_terms.item()) == terms
    bad = []
    for n, prm in m.named_parameters():
        e = rel_err(prm.grad, g64[n])
        allowed = max(REL_TOL, NOISE_MULT * case["noise"].get(n, 0.0))
        if not e <= allowed:
            bad.append((n, e, allowed))
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)
