"""GPU parity: the CUDA path (through the C ABI, via the nn.Module mirror) against
 (a) golden fixtures produced by running the reference in fp64 (tests/golden), and
 (b) the CPU oracle on the same seeded inputs, tensor by tensor.
Tolerance: per tensor max|a-b|/max|b| <= 1e-4 against the fp64 reference, or <= 4x the
fp32 reference's own error vs fp64 for that tensor."""
import ctypes

import numpy as np
import pytest
import torch

from tests.golden_util import NOISE_MULT, REL_TOL, check_parity, load_case, rel_err

pytestmark = pytest.mark.gpu


def _model(case, kind, **kw):
    from pathtracker_models_b200 import InT, ffhgru_hierarchy
    if kind == "int":
        m = InT.InT(dimensions=32, timesteps=8, kernel_size=7, **kw)
    else:
        m = ffhgru_hierarchy.FFhGRU(dimensions=32, timesteps=8, kernel_size=7)
    m.load_state_dict(case["sd"])
    return m.cuda()


def _run(case, m, x=None):
    x = case["x"].float().cuda() if x is None else x
    y = case["y"].float().cuda()
    out, pen = m(x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    return out, loss, grads, pen


CASES = [("int_small", "int", {}), ("int_tile", "int", {}), ("int_cfg1", "int", {}), ("int_64px", "int", {}),
         ("int_tanh_small", "int", {"nl": torch.tanh}), ("int_noinh_small", "int", {"no_inh": True}),
         ("int_lesion_small", "int", {"lesion_alpha": True, "lesion_gamma": True}),
         ("hgru_small", "hgru", {}), ("hgru_tile", "hgru", {})]


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name,kind,kw", CASES)
def test_logits_and_all_grads_vs_fp64_reference(name, kind, kw, conv_mode):
    """both precisions of the path: tcgen05 split-bf16 convolutions (default) and strict-fp32 CUDA cores."""
    from pathtracker_models_b200 import _lib
    case = load_case(name)
    m = _model(case, kind, **kw)
    m.conv_mode = _lib.PTK_CONV_BF16X3 if conv_mode == "bf16x3" else _lib.PTK_CONV_FP32
    out, loss, grads, pen = _run(case, m)
    assert out.shape == (case["shape"][0], 1) and pen.shape == (1,) and pen.device == out.device
    assert abs(loss.item() - case["loss"]) < 1e-5
    bad = check_parity(case, out, grads)
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)
    assert grads["unit1.w"] is None          # dead parameter gets no gradient
    if kw.get("no_inh"):                     # parameters of the skipped inhibition branch: zero gradient
        for n in ("unit1.alpha", "unit1.mu", "unit1.i_w_gate.weight", "unit1.bn.0.weight"):
            assert float(grads[n].abs().max()) == 0.0, n


# The horizons the benchmark is quoted on: BASELINE.json configs[1] (T = 64, 32x32), configs[3] (hGRU, T = 64) and
# configs[2] (T = 128, 64x64), per-GPU batch reduced to what the fp64 reference runs in a minute on the CPU.
# Same gate as the short cases (1e-4 or 4x the fp32 reference's own error), both precisions of the path.
# + the InT variants utils/engine.py builds (tanh, no_inh, lesioned), T = 64, B = 4.
LONG_CASES = [("int_t64", "int", {}), ("hgru_t64", "hgru", {}), ("int_t128_64px", "int", {}),
              ("int_tanh_t64", "int", {"nl": torch.tanh}), ("int_noinh_t64", "int", {"no_inh": True}),
              ("int_lesion_t64", "int", {"lesion_mu": True, "lesion_kappa": True})]


@pytest.mark.parametrize("conv_mode", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name,kind,kw", LONG_CASES)
def test_long_horizon_logits_and_all_grads_vs_fp64_reference(name, kind, kw, conv_mode):
    from pathtracker_models_b200 import _lib
    case = load_case(name)
    m = _model(case, kind, **kw)
    m.conv_mode = _lib.PTK_CONV_BF16X3 if conv_mode == "bf16x3" else _lib.PTK_CONV_FP32
    out, loss, grads, pen = _run(case, m)
    assert abs(loss.item() - case["loss"]) < 1e-5
    bad = check_parity(case, out, grads)
    assert not bad, "\n".join(f"{n}: err {e:.3e} > allowed {a:.3e}" for n, e, a in bad)


# The single-pass bf16 mode (BASELINE cfg 4 "bf16 variants reported separately") is NOT an fp32-parity
# path: operands are rounded to 8 mantissa bits (2^-9 ~ 2e-3 relative per element).  Its error is
# documented and bounded here; it is opt-in (conv_mode=PTK_CONV_BF16) and never the default.
BF16_LOGIT_TOL = 2e-2
BF16_GRAD_TOL = 3e-2     # measured worst over these cases: 1.28e-2 (int_tile), 1.17e-2 (hgru_small); T = 64: 3.7e-3 / 6.3e-3 (profiles/r4q_parity_report.txt)


@pytest.mark.parametrize("name,kind,kw", [c for c in CASES if c[0] in ("int_small", "int_tile", "int_64px", "hgru_small", "hgru_tile", "int_cfg1")])
def test_bf16_single_pass_mode_error_is_bounded(name, kind, kw):
    from pathtracker_models_b200 import _lib
    case = load_case(name)
    m = _model(case, kind, **kw)
    m.conv_mode = _lib.PTK_CONV_BF16
    out, loss, grads, pen = _run(case, m)
    assert rel_err(out, case["logits"]) < BF16_LOGIT_TOL
    worst = max(rel_err(g, case["grads"][n]) for n, g in grads.items()
                if g is not None and n in case["grads"] and float(case["grads"][n].abs().max()) > 0)
    assert worst < BF16_GRAD_TOL, worst
    assert worst > 1e-5          # the mode really ran reduced precision (guards a silent mode mix-up)


def test_per_step_checkpoints_vs_oracle():
    """localises a failure: every I_t, E_t, pre-BN conv output and BN statistic of the CUDA forward."""
    from oracle import int_oracle as O
    from pathtracker_models_b200.functional import int_forward_raw
    from pathtracker_models_b200 import _lib
    case = load_case("int_small")
    m = _model(case, "int")
    m.conv_mode = _lib.PTK_CONV_FP32          # strict mode: per-step tensors agree to ~1e-6
    x = case["x"].float().cuda()
    logits, saved, _, _ = int_forward_raw(x, m.flat_parameters().detach(), m._config(), want_saved=True)
    B, T, H, W = case["shape"]
    S = B * 32 * H * W
    sv = saved.view(torch.float32).cpu()
    off = [0]
    for n in (T + 1, T + 1, T, T, T, T, T):      # ckI, ckE, Y0, Y1, A, Gi, Gex (ptk_api.cu:bind_saved)
        off.append(off[-1] + n * S)
    ckI = sv[off[0]:off[1]].view(T + 1, B, 32, H, W)
    ckE = sv[off[1]:off[2]].view(T + 1, B, 32, H, W)
    Y0 = sv[off[2]:off[3]].view(T, B, 32, H, W)
    Y1 = sv[off[3]:off[4]].view(T, B, 32, H, W)
    st = sv[off[7]:].view(T, 2, 2, 32)
    p = O.unpack(case["sd"])
    with torch.no_grad():
        lo, so = O.forward(case["x"], p)
    for t in range(T):
        assert rel_err(Y0[t], so["Y0"][t]) < 2e-5 or so["Y0"][t].abs().max() == 0, f"Y0[{t}]"
        assert rel_err(st[t, 0, 0], so["m0"][t]) < 2e-5 or t == 0, f"mean0[{t}]"
        assert rel_err(st[t, 0, 1], so["r0"][t]) < 2e-5, f"rstd0[{t}]"
        assert rel_err(ckI[t + 1], so["I"][t + 1]) < 2e-5, f"I[{t + 1}]"
        assert rel_err(Y1[t], so["Y1"][t]) < 2e-5, f"Y1[{t}]"
        assert rel_err(st[t, 1, 0], so["m1"][t]) < 2e-5, f"mean1[{t}]"
        assert rel_err(st[t, 1, 1], so["r1"][t]) < 2e-5, f"rstd1[{t}]"
        assert rel_err(ckE[t + 1], so["E"][t + 1]) < 2e-5, f"E[{t + 1}]"
    assert rel_err(logits, lo) < 1e-5


@pytest.mark.parametrize("name,kind", [("int_small", "int"), ("hgru_small", "hgru")])
def test_testmode_outputs(name, kind):
    case = load_case(name)
    m = _model(case, kind)
    out, states, gates = m(case["x"].float().cuda(), testmode=True)
    B, T, H, W = case["shape"]
    assert states.shape == (B, T, 1, H, W) and gates.shape == (B, T, 32, H, W)
    assert rel_err(out, case["extra"]["tm_logits"]) < REL_TOL
    assert rel_err(states, case["extra"]["tm_states"]) < REL_TOL
    assert rel_err(gates, case["extra"]["tm_gates"]) < REL_TOL


def test_channels_last_3d_input_and_eval_mode():
    """utils/engine.py:222-243 hands over a channels_last_3d-strided tensor; eval() == train()."""
    case = load_case("int_small")
    m = _model(case, "int")
    xc = case["x"].float().cuda()
    xs = xc.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
    assert not xs.is_contiguous()
    o1, _, g1, _ = _run(case, m, xc)
    m.zero_grad(set_to_none=True)
    o2, _, g2, _ = _run(case, m, xs)
    assert torch.equal(o1, o2)
    for n in g1:
        assert (g1[n] is None and g2[n] is None) or torch.equal(g1[n], g2[n]), n
    m.eval()
    with torch.no_grad():
        o3, _ = m(xc)
    assert torch.equal(o1, o3)


def test_determinism_and_input_gradient():
    case = load_case("int_tile")
    m = _model(case, "int")
    x = case["x"].float().cuda().requires_grad_(True)
    outs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        x.grad = None
        out, _ = m(x)
        out.sum().backward()
        outs.append((out.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    for n in outs[0][2]:
        assert torch.equal(outs[0][2][n], outs[1][2][n]), n
    # input gradient (viz_model_att.py:189-196) against oracle autograd in fp64
    from oracle import int_oracle as O
    p = O.unpack(case["sd"])
    xo = case["x"].clone().requires_grad_(True)
    lo, _ = O.forward(xo, p)
    lo.sum().backward()
    assert rel_err(outs[0][1], xo.grad) < REL_TOL


def test_synth_generator_bit_exact_and_well_formed():
    from oracle.synth_oracle import synth
    from pathtracker_models_b200 import synth_pathtracker
    for (B, T, H, W, D, speed, first) in [(5, 16, 32, 32, 14, 1, 0), (3, 8, 64, 64, 0, 4, 1000), (4, 12, 32, 32, 25, 2, 7)]:
        x, y = synth_pathtracker(B, T, H, W, n_distractors=D, speed=speed, seed=1234, first_index=first)
        xo, yo = synth(1234, first, B, T, H, W, D, speed)
        assert np.array_equal(x.cpu().numpy(), xo) and np.array_equal(y.cpu().numpy().ravel(), yo)
        assert set(np.unique(xo)) <= {0.0, 1.0} and x.shape == (B, 3, T, H, W) and y.shape == (B, 1)
        assert (xo[:, 0].reshape(B, T, -1).sum(-1) >= 1).all()          # the target is always drawn
        assert (xo[:, 1].reshape(B, T, -1).sum(-1) == 9).all() and (xo[:, 2].reshape(B, T, -1).sum(-1) == 9).all()
    # disjoint deterministic shards: rank r's shard == slice of the global batch
    xa, _ = synth_pathtracker(8, 8, 32, 32, first_index=0)
    xb, _ = synth_pathtracker(4, 8, 32, 32, first_index=4)
    assert torch.equal(xa[4:], xb)


def test_errors_are_loud():
    from pathtracker_models_b200 import InT, _lib
    m = InT.InT(dimensions=32, timesteps=4, kernel_size=7).cuda()
    with pytest.raises(_lib.PtkError):
        m(torch.rand(2, 3, 4, 20, 20, device="cuda"))          # unsupported frame size
    with pytest.raises(_lib.PtkError):
        m(torch.rand(2, 3, 4, 32, 32))                         # CPU tensor: no fallback
    with pytest.raises(_lib.PtkError):
        InT.InT(dimensions=32, timesteps=4, kernel_size=15).cuda()(torch.rand(1, 3, 2, 32, 32, device="cuda"))
    with pytest.raises(_lib.PtkError, match="readout tile"):
        m(torch.rand(1, 3, 1, 160, 160, device="cuda"))         # frame larger than the readout kernels' shared-memory tile
    d = _lib.PtkDims(1, 1, 32, 32, 32, 7, 0, 0, 0, 0)
    L = _lib.lib()
    rc = L.ptk_int_forward(ctypes.byref(d), None, None, None, None, None, None, None, None, 0, None)
    assert rc == 1 and b"NULL" in L.ptk_last_error()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(3, 32, 32), (2, 64, 64)])
def test_conv7_operator_vs_torch(mode, shape):
    """ptk_conv7 (fp32 FFMA path and 3-pass split-bf16 tcgen05 path) against torch conv2d in fp64:
    forward, conv_transpose (data gradient), accumulate, and the BatchNorm batch statistics."""
    import torch.nn.functional as F
    from pathtracker_models_b200.functional import conv7
    B, H, W = shape
    g = torch.Generator().manual_seed(5 + mode)
    x = torch.randn(B, 32, H, W, generator=g) * torch.rand(B, 32, 1, 1, generator=g) * 3
    w = torch.randn(32, 32, 7, 7, generator=g) * 0.05
    xd, wd = x.double(), w.double()
    tol = 2e-6 if mode == 0 else 3e-5
    y, st = conv7(x.cuda(), w.cuda(), conv_mode=mode, want_stats=True)
    ref = F.conv2d(xd, wd, padding=3)
    assert rel_err(y, ref) < tol
    mean = ref.mean(dim=(0, 2, 3))
    rstd = 1 / torch.sqrt(ref.var(dim=(0, 2, 3), unbiased=False) + 1e-3)
    assert rel_err(st[0], mean) < 1e-4 and rel_err(st[1], rstd) < 1e-5
    yt = conv7(x.cuda(), w.cuda(), conv_mode=mode, transpose=True)
    assert rel_err(yt, F.conv_transpose2d(xd, wd, padding=3)) < tol
    acc = torch.randn(B, 32, H, W, generator=g)
    ya = conv7(x.cuda(), w.cuda(), conv_mode=mode, out=acc.clone().cuda())
    assert rel_err(ya, ref + acc.double()) < tol
    y2 = conv7(x.cuda(), w.cuda(), conv_mode=mode)
    assert torch.equal(y, y2)                       # deterministic


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("shape", [(1, 1, 16, 32), (3, 3, 16, 64), (2, 2, 64, 64), (5, 7, 32, 32), (2, 64, 32, 32)])
def test_wgrad7_operator_vs_torch(mode, shape):
    """ptk_wgrad7 (ptk_wgrad_ts.cu: the 7x7 weight gradient summed over a whole sequence in one launch, accumulators and
    the dY operand in TMEM) against torch autograd in fp64: single tile, ragged batch, frames wider / taller than a tile
    (real halo columns), more work items than CTAs.  mode 1 = 3-pass split-bf16 (fp32-parity), 2 = single-pass bf16."""
    import torch.nn.functional as F
    from pathtracker_models_b200.functional import wgrad7
    T, B, H, W = shape
    g = torch.Generator().manual_seed(T * 100 + B + mode)
    inp = torch.randn(T, B, 32, H, W, generator=g) * torch.rand(1, 1, 32, 1, 1, generator=g) * 2
    dy = torch.randn(T, B, 32, H, W, generator=g) * 0.1
    w = torch.zeros(32, 32, 7, 7, dtype=torch.float64, device="cuda", requires_grad=True)
    tot = sum((F.conv2d(inp[t].double().cuda(), w, padding=3) * dy[t].double().cuda()).sum() for t in range(T))
    tot.backward()
    got = wgrad7(inp.cuda(), dy.cuda(), conv_mode=mode)
    e = rel_err(got, w.grad)
    assert e < (3e-5 if mode == 1 else 2e-2), e
    if mode == 2:
        assert e > 1e-5                                         # really the single-pass mode
    assert torch.equal(got, wgrad7(inp.cuda(), dy.cuda(), conv_mode=mode))      # deterministic


def test_device_prefetcher_copies_every_batch():
    from pathtracker_models_b200.data import DevicePrefetcher
    host = [(torch.randn(4, 8).pin_memory(), torch.randn(4, 1).pin_memory()) for _ in range(4)]
    n = 0
    for (hx, hy), (x, y) in zip(host, DevicePrefetcher(iter(host), "cuda")):
        assert x.is_cuda and torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)
        n += 1
    assert n == 4


def test_programmatic_dependent_launch_changes_no_bits(tmp_path):
    """the PDL chain (ptk_common.cuh) must be pure scheduling: the same step with the launch attribute disabled
    (PTK_NO_PDL=1, read once per process) is bit-identical; a missing griddepcontrol.wait would show up here."""
    import os
    import subprocess
    import sys
    script = r'''
import sys, torch
sys.path.insert(0, %r)
from tests.golden_util import load_case
from pathtracker_models_b200 import InT
case = load_case("int_tile")
m = InT.InT(dimensions=32, timesteps=8, kernel_size=7); m.load_state_dict(case["sd"]); m = m.cuda()
x = case["x"].float().cuda(); y = case["y"].float().cuda()
outs = []
for _ in range(3):
    m.zero_grad(set_to_none=True)
    out, _ = m(x)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
    outs.append(torch.cat([out.reshape(-1)] + [p.grad.reshape(-1) for p in m.parameters() if p.grad is not None]).cpu())
assert all(torch.equal(outs[0], o) for o in outs[1:])
torch.save(outs[0], sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("pdl", {}), ("nopdl", {"PTK_NO_PDL": "1"})):
        f = str(tmp_path / (tag + ".pt"))
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", script, f], env=e, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = torch.load(f)
    assert torch.equal(res["pdl"], res["nopdl"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (nn.DataParallel, mainclean.py:132-134)")
def test_nn_dataparallel_two_devices_matches_per_replica_runs():
    """the reference's own multi-GPU mode: replicas called concurrently from Python threads, BatchNorm statistics
    per replica, gradients summed onto device 0 (SURVEY 8b threading / 8e)."""
    case = load_case("int_small")
    m = _model(case, "int")
    x = torch.cat([case["x"].float(), case["x"].float().flip(0)], 0).cuda()       # batch 2B -> B per device
    y = torch.cat([case["y"].float(), case["y"].float().flip(0)], 0).cuda()
    B = x.shape[0] // 2
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    out, pen = dp(x)
    assert out.shape == (2 * B, 1) and pen.shape == (2,)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    g_dp = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    outs = []
    for h in range(2):       # the same halves on one device, each weighted 1/2 in the full-batch mean
        o, _ = m(x[h * B:(h + 1) * B])
        (0.5 * torch.nn.functional.binary_cross_entropy_with_logits(o, y[h * B:(h + 1) * B])).backward()
        outs.append(o.detach())
    assert rel_err(out, torch.cat(outs, 0)) < 1e-6
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert rel_err(g_dp[n], p.grad) < 1e-5, n


def test_saliency_through_testmode_forward():
    """viz_model_att.py:189-196: imgs.requires_grad = True; output, states, gates = model(imgs, testmode=True);
    score(output).backward(); imgs.grad -- the testmode logits must stay differentiable."""
    case = load_case("int_small")
    m = _model(case, "int")
    x1 = case["x"].float().cuda().requires_grad_(True)
    out1, states, gates = m(x1, testmode=True)
    assert not states.requires_grad and not gates.requires_grad
    torch.nn.functional.mse_loss(out1.flatten(), torch.zeros_like(out1.flatten())).backward()
    x2 = case["x"].float().cuda().requires_grad_(True)
    out2, _ = m(x2)
    torch.nn.functional.mse_loss(out2.flatten(), torch.zeros_like(out2.flatten())).backward()
    assert torch.equal(out1, out2)
    assert x1.grad is not None and float(x1.grad.abs().max()) > 0
    assert rel_err(x1.grad, x2.grad) < 1e-6
    assert rel_err(gates, case["extra"]["tm_gates"]) < REL_TOL


@pytest.mark.parametrize("kind,B,T,H,W", [("int", 1, 1, 16, 32), ("int", 5, 2, 16, 64), ("hgru", 3, 3, 48, 32),
                                           ("int", 2, 2, 64, 64), ("int", 96, 1, 32, 32), ("int", 1, 2, 96, 96)])
def test_edge_shapes_vs_fp64_oracle(kind, B, T, H, W):
    """minimum tile (16x32), single sample / single frame, odd batch, non-square frames, a batch larger than the SM
    count: CUDA path vs the fp64 oracle on the same seeded inputs (the oracle is pinned to the reference on CPU).
    Note T = 1 exercises the t = 0 edge case: conv input is all-zero, BatchNorm variance 0."""
    from oracle import int_oracle as O
    from pathtracker_models_b200 import InT, ffhgru_hierarchy
    torch.manual_seed(B * 1000 + T * 100 + H + W)
    ctor = InT.InT if kind == "int" else ffhgru_hierarchy.FFhGRU
    m = ctor(dimensions=32, timesteps=T, kernel_size=7).cuda()
    x = torch.rand(B, 3, T, H, W, device="cuda")
    y = (torch.rand(B, 1, device="cuda") > 0.5).float()
    out, _ = m(x)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y).backward()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    p = O.unpack(sd)
    for v in p.values():
        v.requires_grad_(True)
    lo, _ = O.forward(x.double().cpu(), p, variant=kind, keep=False)
    torch.nn.functional.binary_cross_entropy_with_logits(lo, y.double().cpu()).backward()
    go = O.grad_to_state_keys({k: v.grad for k, v in p.items() if v.grad is not None}, sd)
    assert rel_err(out, lo) < REL_TOL
    # the golden gate: 1e-4, or 4 x the error of the reference algorithm itself run in fp32 on this input
    p32 = O.unpack(sd, dtype=torch.float32)
    for v in p32.values():
        v.requires_grad_(True)
    l32, _ = O.forward(x.float().cpu(), p32, variant=kind, keep=False)
    torch.nn.functional.binary_cross_entropy_with_logits(l32, y.float().cpu()).backward()
    g32 = O.grad_to_state_keys({k: v.grad for k, v in p32.items() if v.grad is not None}, sd)
    bad = []
    for n, prm in m.named_parameters():
        if prm.grad is not None and float(go[n].abs().max()) > 0:
            e, allowed = rel_err(prm.grad, go[n]), max(REL_TOL, NOISE_MULT * rel_err(g32[n], go[n]))
            if not e <= allowed:
                bad.append((n, e, allowed))
    assert not bad, bad
