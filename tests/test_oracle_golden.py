"""CPU: pin the oracle restatement against fixtures produced by running the reference."""
import pytest
import torch

from oracle import int_oracle as O
from tests.golden_util import load_case, rel_err

CASES = [("int_small", "int", "softplus", False), ("int_tile", "int", "softplus", False),
         ("int_64px", "int", "softplus", False), ("int_tanh_small", "int", "tanh", False),
         ("int_noinh_small", "int", "softplus", True), ("int_lesion_small", "int", "softplus", False),
         ("hgru_small", "hgru", "softplus", False), ("hgru_tile", "hgru", "softplus", False)]


def _autograd(case, variant, act, no_inh):
    p = O.unpack(case["sd"])
    for v in p.values():
        v.requires_grad_(True)
    logits, saved = O.forward(case["x"], p, variant, act, no_inh)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, case["y"])
    loss.backward()
    g = O.grad_to_state_keys({k: v.grad for k, v in p.items() if v.grad is not None}, case["sd"])
    return logits.detach(), loss.item(), g


@pytest.mark.parametrize("name,variant,act,no_inh", CASES)
def test_oracle_forward_and_autograd_match_reference_fp64(name, variant, act, no_inh):
    case = load_case(name)
    logits, loss, g = _autograd(case, variant, act, no_inh)
    assert rel_err(logits, case["logits"]) < 1e-10
    assert abs(loss - case["loss"]) < 1e-10
    for n, g_ref in case["grads"].items():
        big = g_ref.numel() > 4096           # stored as float32 in the fixture
        assert rel_err(g[n], g_ref) < (2e-7 if big else 1e-9), n


@pytest.mark.parametrize("name,variant,act", [(c[0], c[1], c[2]) for c in CASES if not c[3]])
def test_oracle_manual_backward_matches_reference_fp64(name, variant, act):
    """the hand-derived reverse-time loop (what the CUDA kernels implement) vs the reference grads."""
    case = load_case(name)
    p = O.unpack(case["sd"])
    with torch.no_grad():
        logits, saved = O.forward(case["x"], p, variant, act)
        dlogits = O.bce_with_logits_grad(logits, case["y"])
        g = O.grad_to_state_keys(O.backward_manual(case["x"], p, saved, dlogits, variant, act), case["sd"])
    for n, g_ref in case["grads"].items():
        big = g_ref.numel() > 4096
        assert rel_err(g[n], g_ref) < (2e-7 if big else 1e-9), n


# the horizons the benchmark is quoted on (BASELINE.json configs[1]-[3]): T = 64 at 32x32 (B = 8), hGRU likewise,
# T = 128 at 64x64 (B = 2).  Fixtures = the unmodified reference in fp64 (oracle/gen_golden.py long).
# + the InT variants utils/engine.py builds, at T = 64 (B = 4): tanh, no_inh, lesioned (oracle/gen_golden.py long_variants).
LONG_CASES = [("int_t64", "int", "softplus", False), ("hgru_t64", "hgru", "softplus", False), ("int_t128_64px", "int", "softplus", False),
              ("int_tanh_t64", "int", "tanh", False), ("int_noinh_t64", "int", "softplus", True),
              ("int_lesion_t64", "int", "softplus", False)]


@pytest.mark.parametrize("name,variant,act,no_inh", LONG_CASES)
def test_oracle_long_horizon_matches_reference_fp64(name, variant, act, no_inh):
    case = load_case(name)
    logits, loss, g = _autograd(case, variant, act, no_inh)
    assert rel_err(logits, case["logits"]) < 1e-10
    assert abs(loss - case["loss"]) < 1e-10
    for n, g_ref in case["grads"].items():
        assert rel_err(g[n], g_ref) < (2e-7 if g_ref.numel() > 4096 else 1e-8), n


def test_oracle_testmode_outputs():
    for name, variant in (("int_small", "int"), ("hgru_small", "hgru")):
        case = load_case(name)
        p = O.unpack(case["sd"])
        with torch.no_grad():
            logits, saved = O.forward(case["x"], p, variant, testmode=True)
        assert rel_err(torch.stack(saved["states"], 1), case["extra"]["tm_states"]) < 1e-6
        assert rel_err(torch.stack(saved["gates"], 1), case["extra"]["tm_gates"]) < 1e-6
        assert rel_err(logits, case["extra"]["tm_logits"]) < 1e-10


def test_oracle_cfg1_logits():
    case = load_case("int_cfg1")
    p = O.unpack(case["sd"])
    with torch.no_grad():
        logits, _ = O.forward(case["x"], p, keep=False)
    assert rel_err(logits, case["logits"]) < 1e-10


@pytest.mark.parametrize("name,variant", [("int_small", "int"), ("int_tile", "int"), ("hgru_small", "hgru")])
def test_library_operator_restatement_matches_reference_fp64(name, variant):
    """oracle/int_eager.py (the reference's operator sequence, timed by `bench.py --impl reference --ref-device cuda`)
    against the fixtures produced by the reference itself: logits and every parameter gradient."""
    from oracle import int_eager
    case = load_case(name)
    p = O.unpack(case["sd"])
    for v in p.values():
        v.requires_grad_(True)
    logits = int_eager.forward(case["x"], p, variant)
    torch.nn.functional.binary_cross_entropy_with_logits(logits, case["y"]).backward()
    g = O.grad_to_state_keys({k: v.grad for k, v in p.items() if v.grad is not None}, case["sd"])
    assert rel_err(logits.detach(), case["logits"]) < 1e-10
    for n, g_ref in case["grads"].items():
        assert rel_err(g[n], g_ref) < (2e-7 if g_ref.numel() > 4096 else 1e-9), n
