"""CPU: the drop-in boundary -- constructors, state_dict names/shapes, init RNG order, the C ABI
library loads and exports every symbol include/ptk_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pathtracker_models_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ptk_b200.h")).read()
    declared = set(re.findall(r"PTK_API\s+[\w\s\*]+?\b(ptk_\w+)\s*\(", hdr))
    assert declared and declared == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().ptk_version() == 100


def test_rbp_checkpoints_are_constant_memory():
    """recurrent back-propagation (models/convlstm.py:124-135) keeps no per-step state: with ckpt_ring = 1 the checkpoint
    buffer of the ConvLSTM does not grow with the number of time steps; the BPTT buffer does."""
    from pathtracker_models_b200 import _lib
    L = _lib.lib()
    sb = lambda T, ring: L.ptk_convlstm_saved_bytes(ctypes.byref(_lib.PtkLstmDims(8, T, 32, 32, 15, 1, ring)))
    assert sb(4, 1) == sb(64, 1) == sb(1000, 1) > 0
    assert sb(64, 0) > 10 * sb(64, 1)
    per_step = (25 + 25 + 100) * 8 * 32 * 32 * 4            # h_t, c_t, gates_t
    assert sb(65, 0) - sb(64, 0) == per_step


def test_param_layout_matches_module():
    from pathtracker_models_b200 import InT, _lib
    from pathtracker_models_b200._recurrent import PARAM_ORDER
    m = InT.InT(dimensions=32, timesteps=4, kernel_size=7)
    off = _lib.param_offsets(_lib.PtkDims(1, 1, 32, 32, 32, 7, 0, 0, 0, 0))
    named = dict(m.named_parameters())
    assert len(PARAM_ORDER) == _lib.PTK_INT_NPARAM
    assert [off[i + 1] - off[i] for i in range(30)] == [named[n].numel() for n in PARAM_ORDER]
    assert off[-1] == m.flat_parameters().numel() == 107190 - 32      # minus the dead unit1.w


def test_dims_validation_without_gpu():
    from pathtracker_models_b200 import _lib
    L = _lib.lib()
    assert L.ptk_int_workspace_bytes(ctypes.byref(_lib.PtkDims(64, 64, 32, 32, 32, 7, 0, 0, 0, 0))) > 0
    assert L.ptk_int_saved_bytes(ctypes.byref(_lib.PtkDims(64, 64, 32, 32, 32, 7, 0, 0, 0, 0))) == \
        ((2 * 65 + 5 * 64) * 64 * 32 * 1024 + 64 * 128) * 4
    assert L.ptk_int_workspace_bytes(ctypes.byref(_lib.PtkDims(2, 4, 20, 20, 32, 7, 0, 0, 0, 0))) == 0
    assert b"unsupported shape" in L.ptk_last_error()
    assert L.ptk_int_workspace_bytes(ctypes.byref(_lib.PtkDims(2, 4, 32, 32, 32, 15, 0, 0, 0, 0))) == 0
    assert b"kernel_size=7" in L.ptk_last_error()


@pytest.mark.parametrize("which", ["int", "hgru"])
def test_constructor_is_bit_identical_to_reference_init(which):
    """torch.manual_seed(0); New(...) consumes the RNG exactly like the reference constructor
    (fixture written by oracle/gen_golden.py from the reference's own state_dict)."""
    from pathtracker_models_b200 import InT, ffhgru_hierarchy
    z = np.load(os.path.join(GOLDEN, f"{which}_fingerprint.npz"))
    torch.manual_seed(0)
    m = InT.InT(dimensions=32, timesteps=32, kernel_size=7) if which == "int" else \
        ffhgru_hierarchy.FFhGRU(dimensions=32, timesteps=32, kernel_size=7)
    sd = m.state_dict()
    keys = [k for k in sd if not (which == "int" and k == "unit1.w")]
    assert keys == list(z["init_keys"])
    sums = np.array([sd[k].double().sum().item() for k in keys])
    assert np.array_equal(sums, z["init_sums"])


def test_state_dict_round_trip_and_variants():
    from pathtracker_models_b200 import InT, ffhgru_hierarchy
    case = load_case("int_small")
    m = InT.InT(dimensions=32, timesteps=8, kernel_size=7)
    m.load_state_dict(case["sd"])                       # a reference checkpoint loads, incl. dead unit1.w
    assert set(m.state_dict().keys()) == set(case["sd"].keys()) and len(m.state_dict()) == 31
    for k, v in m.state_dict().items():
        assert v.shape == case["sd"][k].shape, k
    h = ffhgru_hierarchy.FFhGRU(dimensions=32, timesteps=8, kernel_size=7)
    h.load_state_dict(load_case("hgru_small")["sd"])
    assert len(h.state_dict()) == 33                    # + unused BatchNorm3d (ffhgru_hierarchy.py:186)
    les = InT.InT(32, 8, 7, lesion_alpha=True, lesion_gamma=True)
    assert not les.unit1.alpha.requires_grad and not les.unit1.gamma.requires_grad
    assert float(les.unit1.alpha.mean()) == 1.0          # frozen at init, not zeroed (InT.py:132-143)
    ni = InT.InT(32, 8, 7, no_inh=True)
    assert "unit1.w_inh" not in ni.state_dict()
    with pytest.raises(NotImplementedError):
        InT.InT(32, 8, 7, nl=torch.relu)._config()
    # DataParallel-style key prefixing (earlystopping.py:55 saves module.* keys)
    dp_sd = {"module." + k: v for k, v in m.state_dict().items()}
    m2 = InT.InT(dimensions=32, timesteps=8, kernel_size=7)
    m2.load_state_dict({k[len("module."):]: v for k, v in dp_sd.items()})


def test_no_cpu_fallback():
    from pathtracker_models_b200 import InT, _lib
    m = InT.InT(dimensions=32, timesteps=4, kernel_size=7)
    with pytest.raises(_lib.PtkError):
        m(torch.rand(1, 3, 2, 32, 32))


def test_synth_oracle_is_deterministic_and_binary():
    from oracle.synth_oracle import synth
    x, y = synth(1234, 0, 3, 8, 32, 32, 14, 1)
    x2, y2 = synth(1234, 0, 3, 8, 32, 32, 14, 1)
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and set(np.unique(x)) <= {0.0, 1.0}
    xs, _ = synth(1234, 1, 2, 8, 32, 32, 14, 1)
    assert np.array_equal(x[1:], xs)


def test_device_prefetcher_preserves_order_and_count():
    from pathtracker_models_b200.data import DevicePrefetcher
    batches = [(torch.full((2, 3), float(i)), torch.tensor([float(i)])) for i in range(5)]
    got = list(DevicePrefetcher(iter(batches), "cpu"))
    assert len(got) == 5
    for i, (x, y) in enumerate(got):
        assert float(x[0, 0]) == i and float(y[0]) == i
    assert list(DevicePrefetcher(iter([]), "cpu")) == []


def test_gru_flat_layout_covers_all_fourteen_tensors():
    """ptk_gru_param_offsets (include/ptk_b200.h) vs kys.GRU.flat_parameters: stem, cell and readout head in one vector"""
    import ctypes
    from pathtracker_models_b200 import _lib, kys
    m = kys.GRU(dimensions=64, timesteps=4, kernel_size=7)
    off = (ctypes.c_int64 * 15)()
    dims = _lib.PtkGruDims(2, 4, 32, 32, 64, 7, _lib.PTK_CONV_FP32)
    _lib.check(_lib.lib().ptk_gru_param_offsets(ctypes.byref(dims), off), "ptk_gru_param_offsets")
    u = m.unit1
    ts = [m.preproc.weight, m.preproc.bias, u.conv_reset.weight, u.conv_reset.bias, u.conv_update.weight, u.conv_update.bias,
          u.conv_state_new.weight, u.conv_state_new.bias, m.readout_conv.weight, m.readout_conv.bias, m.target_conv.weight,
          m.target_conv.bias, m.readout_dense.weight, m.readout_dense.bias]
    assert [off[i + 1] - off[i] for i in range(14)] == [t.numel() for t in ts]
    flat = m.flat_parameters()
    assert flat.numel() == off[14]
    assert torch.equal(flat[off[10]:off[11]], m.target_conv.weight.detach().reshape(-1))
    # every trainable tensor except the reference's unused BatchNorm3d (models/kys.py:80) travels in the flat vector
    assert sum(p.numel() for n, p in m.named_parameters() if not n.startswith("bn.")) == off[14]
