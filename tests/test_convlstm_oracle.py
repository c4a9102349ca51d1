"""CPU: pin the ConvLSTM oracle against fixtures produced by running the reference (models/convlstm.py)."""
import numpy as np
import pytest

from tests.convlstm_util import load_convlstm_case, oracle_run
from tests.golden_util import rel_err


@pytest.mark.parametrize("name", ["convlstm_k7_small", "convlstm_k15_small"])
def test_convlstm_oracle_matches_reference_fp64(name):
    case = load_convlstm_case(name)
    out, loss, pen, grads = oracle_run(case)
    assert rel_err(out, case["output"]) < 1e-10
    assert abs(loss - case["loss"]) < 1e-10
    assert rel_err(pen, case["penalty"]) < 1e-9
    for n, g in case["grads"].items():                      # full gradients (filt 7 case)
        assert rel_err(grads[n], g) < (2e-7 if g.numel() > 4096 else 1e-9), n
    prs = np.random.RandomState(case["proj_seed"])
    for n in case["proj"]:                                  # seeded projections (filt 15 case), same draw order
        g = grads[n].numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert np.allclose(v @ g, case["proj"][n], rtol=1e-8, atol=1e-12 * max(1.0, case["gmax"][n])), n
        assert abs(np.abs(g).max() - case["gmax"][n]) <= 1e-9 * case["gmax"][n], n


@pytest.mark.parametrize("name", ["convlstm_k7_rbp", "convlstm_k7_rbp_trunc"])
def test_rbp_oracle_matches_reference_fp64(name):
    """grad_method='rbp' (dummyhgru Neumann series, models/convlstm.py:9-54,124-135): fixture = the reference itself."""
    import torch
    from oracle import convlstm_oracle as CO
    import numpy as _np
    import os
    from tests.golden_util import GOLDEN
    case = load_convlstm_case(name)
    z = _np.load(os.path.join(GOLDEN, name + ".npz"))
    num_iter, calls = int(z["num_iter"]), int(z["rbp_grad_calls"])
    p = {k: v.double().clone().requires_grad_(True) for k, v in case["sd"].items()}
    out, loss, grads, terms = CO.rbp_grads(case["x"], p, case["shape"][1], case["target"], num_iter)
    assert rel_err(out, case["output"]) < 1e-10 and abs(loss - case["loss"]) < 1e-10
    assert terms in (calls, calls - 1) and terms <= num_iter     # the reference tried `calls` terms (the last may be rejected)
    for n, g in case["grads"].items():
        assert rel_err(grads[n], g) < (2e-7 if g.numel() > 4096 else 1e-9), n


# (convlstm_k15_jp_t64, the same at timesteps = 64, takes 100 s of CPU double back-propagation: the GPU suite holds the oracle to
# that fixture's projections instead, evaluating it in fp64 on the device)
@pytest.mark.parametrize("name", ["convlstm_k7_jp", "convlstm_k7_jp_t2", "convlstm_k15_jp", "convlstm_k7_rbp_jp"])
def test_jacobian_penalty_in_graph_oracle_matches_reference_fp64(name):
    """jacobian_penalty=True (create_graph=True, models/convlstm.py:157-160): gradients of loss + 10 * penalty.mean()
    (mainclean.py:192-196).  Fixture = the reference itself in fp64, both gradient methods."""
    import os
    import numpy as _np
    from oracle import convlstm_oracle as CO
    from tests.golden_util import GOLDEN
    case = load_convlstm_case(name)
    z = _np.load(os.path.join(GOLDEN, name + ".npz"))
    method = "rbp" if "rbp" in name else "bptt"
    p = {k: v.double().clone().requires_grad_(True) for k, v in case["sd"].items()}
    out, loss, pen, grads = CO.jp_total_grads(case["x"], p, case["shape"][1], case["target"], grad_method=method,
                                              num_iter=int(z["num_iter"]))
    assert rel_err(out, case["output"]) < 1e-10 and abs(loss - case["loss"]) < 1e-10
    assert rel_err(pen, case["penalty"]) < 1e-9
    for n, g in case["grads"].items():
        assert rel_err(grads[n], g) < (2e-7 if g.numel() > 4096 else 1e-9), n
    prs = _np.random.RandomState(case["proj_seed"])
    for n in case["proj"]:
        g = grads[n].detach().numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert _np.allclose(v @ g, case["proj"][n], rtol=1e-8, atol=1e-12 * max(1.0, case["gmax"][n])), n
    if case["grads"]:      # the penalty term must matter in the fixture: the loss-only gradient is measurably different
        if method == "bptt":
            _, _, _, g0 = oracle_run(case)
        else:
            p0 = {k: v.double().clone().requires_grad_(True) for k, v in case["sd"].items()}
            _, _, g0, _ = CO.rbp_grads(case["x"], p0, case["shape"][1], case["target"], int(z["num_iter"]))
        assert max(rel_err(g0[n], g) for n, g in case["grads"].items() if "unit1.Wh" in n) > 1e-3


def test_rbp_oracle_long_horizon_matches_reference_fp64():
    """grad_method='rbp' at configs[4]'s own horizon (filt 15, timesteps = 64): the reference's gradients through seeded projections"""
    import os
    import numpy as _np
    from oracle import convlstm_oracle as CO
    from tests.golden_util import GOLDEN
    case = load_convlstm_case("convlstm_k15_rbp_t64")
    z = _np.load(os.path.join(GOLDEN, "convlstm_k15_rbp_t64.npz"))
    num_iter, calls = int(z["num_iter"]), int(z["rbp_grad_calls"])
    p = {k: v.double().clone().requires_grad_(True) for k, v in case["sd"].items()}
    out, loss, grads, terms = CO.rbp_grads(case["x"], p, case["shape"][1], case["target"], num_iter)
    assert rel_err(out, case["output"]) < 1e-10 and abs(loss - case["loss"]) < 1e-10
    assert terms in (calls, calls - 1) and terms <= num_iter
    prs = _np.random.RandomState(case["proj_seed"])
    for n in case["proj"]:
        g = grads[n].detach().numpy().ravel()
        v = prs.normal(0, 1, (4, g.size))
        assert _np.allclose(v @ g, case["proj"][n], rtol=1e-8, atol=1e-12 * max(1.0, case["gmax"][n])), n
