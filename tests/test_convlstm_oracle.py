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
