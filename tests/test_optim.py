"""The step either side of the path: BCE-with-logits + metrics, fused Adam."""
import numpy as np
import pytest
import torch

from oracle import optim_oracle as O


def test_oracle_bce_scores_against_torch_and_a_hand_case():
    rs = np.random.RandomState(0)
    z = rs.randn(37) * 2
    y = (rs.rand(37) > 0.5).astype(np.float64)
    r = O.bce_scores(z, y)
    zt = torch.tensor(z, requires_grad=True)
    lt = torch.nn.functional.binary_cross_entropy_with_logits(zt, torch.tensor(y))
    lt.backward()
    assert abs(r["loss"] - lt.item()) < 1e-12
    assert np.abs(r["dlogits"] - zt.grad.numpy()).max() < 1e-14
    # hand case (utils/misc_functions.py:12-45): logits > 0.5 are positive predictions
    h = O.bce_scores([2.0, 0.4, 0.6, -1.0], [1, 1, 0, 0])
    assert h["bal_acc"] == 50.0 and h["precision"] == 0.5 and h["recall"] == 0.25 and abs(h["f1"] - 2 / 6) < 1e-15
    none = O.bce_scores([-1.0, -2.0], [1, 0])      # no positive prediction: tpfp -> 1e-6
    assert none["precision"] == 0.0 and none["bal_acc"] == 50.0


def test_oracle_adam_reference_matches_closed_form_first_step():
    p = [torch.tensor([1.0, -2.0])]
    g = [[torch.tensor([0.5, -0.25])]]
    out = O.adam_reference(p, g, lr=0.1)[0]
    # first Adam step moves every coordinate by lr * sign(g) (up to eps)
    assert torch.allclose(out, torch.tensor([0.9, -1.9], dtype=torch.float64), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 64, 300])
def test_bce_scores_kernel_vs_oracle(B):
    from pathtracker_models_b200.optim import bce_with_logits_and_scores
    rs = np.random.RandomState(B)
    z = (rs.randn(B, 1) * 2).astype(np.float32)
    y = (rs.rand(B, 1) > 0.5).astype(np.float32)
    zt = torch.tensor(z, device="cuda", requires_grad=True)
    loss, scores = bce_with_logits_and_scores(zt, torch.tensor(y, device="cuda"))
    (3.0 * loss).backward()
    r = O.bce_scores(z.astype(np.float64), y)
    assert abs(loss.item() - r["loss"]) < 2e-6 * max(1.0, abs(r["loss"]))
    got = scores.cpu().numpy()
    want = np.array([r["bal_acc"], r["precision"], r["recall"], r["f1"]])
    assert np.abs(got - want).max() < 1e-4, (got, want)
    assert np.abs(zt.grad.cpu().numpy().reshape(-1) - 3.0 * r["dlogits"]).max() < 1e-6
    assert zt.grad.shape == zt.shape


@pytest.mark.gpu
def test_fused_adam_vs_torch_adam_and_module_contract():
    from pathtracker_models_b200 import InT
    from pathtracker_models_b200.optim import FusedAdam
    torch.manual_seed(0)
    m = InT.InT(dimensions=32, timesteps=4, kernel_size=7, lesion_alpha=True).cuda()
    names = [n for n, _ in m.named_parameters()]
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    trainable = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
    opt = FusedAdam(m.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.StepLR(opt, 2, gamma=0.5)
    assert all(torch.equal(before[n], p.detach()) for n, p in m.named_parameters())   # re-homing keeps values
    steps = []
    g = torch.Generator().manual_seed(1)
    for s in range(5):
        grads = []
        for n, p in trainable:
            grads.append(None if n == "unit1.w" or (s == 2 and n == "unit1.mu") else torch.randn(p.shape, generator=g))
        steps.append(grads)
    # reference: torch.optim.Adam + StepLR in fp64 on CPU
    ref_ps = [torch.nn.Parameter(before[n].double().cpu()) for n, _ in trainable]
    ref = torch.optim.Adam(ref_ps, lr=1e-2)
    rsched = torch.optim.lr_scheduler.StepLR(ref, 2, gamma=0.5)
    for grads in steps:
        for (n, p), rp, gr in zip(trainable, ref_ps, grads):
            p.grad = None if gr is None else gr.cuda()
            rp.grad = None if gr is None else gr.double()
        opt.step(zero_grad=True)
        sched.step()
        ref.step(); ref.zero_grad(); rsched.step()
        assert all(p.grad is None for _, p in trainable)
    for (n, p), rp in zip(trainable, ref_ps):
        err = (p.detach().double().cpu() - rp.detach()).abs().max().item()
        assert err < 2e-6 * max(1.0, rp.detach().abs().max().item()), (n, err)
    assert [n for n, _ in m.named_parameters()] == names and set(m.state_dict().keys()) >= set(names)
    assert torch.equal(m.state_dict()["unit1.alpha"], before["unit1.alpha"])          # frozen (lesioned) stays put
    # the model still runs on the re-homed parameters
    x = torch.rand(2, 3, 4, 32, 32, device="cuda")
    out, _ = m(x)
    out.sum().backward()
    assert torch.isfinite(out).all()


@pytest.mark.gpu
def test_fused_adam_state_dict_round_trip_and_guards():
    """optimizer.state_dict() / load_state_dict() carry the fused moments and step counts (a resumed run must not restart Adam's
    bias correction); add_param_group after construction and a re-assigned p.data are refused instead of silently ignored."""
    from pathtracker_models_b200.optim import FusedAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(5, 3, device="cuda")), torch.nn.Parameter(torch.randn(7, device="cuda"))]
    opt = FusedAdam(ps, lr=1e-2)
    g = torch.Generator().manual_seed(2)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in ps] for _ in range(4)]
    for k in range(2):
        for p, gr in zip(ps, grads[k]):
            p.grad = gr.clone()
        opt.step(zero_grad=True)
    sd = opt.state_dict()
    snap = [p.detach().clone() for p in ps]
    for k in range(2, 4):                      # continue two more steps: the reference trajectory
        for p, gr in zip(ps, grads[k]):
            p.grad = gr.clone()
        opt.step(zero_grad=True)
    want = [p.detach().clone() for p in ps]
    ps2 = [torch.nn.Parameter(s.clone()) for s in snap]
    opt2 = FusedAdam(ps2, lr=1e-2)
    opt2.load_state_dict(sd)
    for k in range(2, 4):
        for p, gr in zip(ps2, grads[k]):
            p.grad = gr.clone()
        opt2.step(zero_grad=True)
    for a, b in zip(ps2, want):
        assert torch.equal(a.detach(), b)
    with pytest.raises(ValueError):
        opt2.add_param_group({"params": [torch.nn.Parameter(torch.zeros(2, device="cuda"))]})
    ps2[0].data = torch.zeros(5, 3, device="cuda")
    ps2[0].grad = torch.ones(5, 3, device="cuda")
    with pytest.raises(RuntimeError, match="re-assigned"):
        opt2.step()
