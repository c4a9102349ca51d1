"""CPU oracle for the ConvGRU baseline (models/kys.py).  TEST INFRASTRUCTURE ONLY.

Restates `ConvGRUCell.forward` (models/kys.py:36-48) and `GRU.forward` (:104-135) with torch CPU ops in any float
dtype; gradients come from autograd on this restatement.  Pinned against outputs of the reference itself
(oracle/gen_golden.py gru -> tests/golden/gru_*.npz, checked by tests/test_gru_oracle.py)."""
import torch
import torch.nn.functional as F


def cell(x_t, h, p, pad):
    zr = torch.cat([x_t, h], 1)                                                                        # kys.py:37
    r = torch.sigmoid(F.conv2d(zr, p["unit1.conv_reset.weight"], p["unit1.conv_reset.bias"], padding=pad))    # :39
    u = torch.sigmoid(F.conv2d(zr, p["unit1.conv_update.weight"], p["unit1.conv_update.bias"], padding=pad))  # :40
    z2 = torch.cat([x_t, r * h], 1)                                                                    # :42
    n = torch.tanh(F.conv2d(z2, p["unit1.conv_state_new.weight"], p["unit1.conv_state_new.bias"], padding=pad))  # :43
    return (1.0 - u) * h + u * n, r                                                                    # :45


def forward(x, p, testmode=False):
    """x [B,3,T,H,W]; p: state_dict-keyed tensors.  Returns logits [B,1] (and states, gates in testmode)."""
    C = p["preproc.bias"].numel()
    pad = p["unit1.conv_reset.weight"].shape[-1] // 2
    xbn = F.softplus(F.conv3d(x, p["preproc.weight"], p["preproc.bias"]))                              # :106-108
    B, _, T, H, W = x.shape
    h = x.new_zeros(B, C, H, W)                                                                        # :112
    states, gates = [], []
    for t in range(T):                                                                                 # :117
        h, r = cell(xbn[:, :, t], h, p, pad)
        if testmode:
            gates.append(r)
            states.append(F.conv2d(h, p["readout_conv.weight"], p["readout_conv.bias"]))
    out = torch.cat([F.conv2d(h, p["readout_conv.weight"], p["readout_conv.bias"]), x[:, 2, 0][:, None]], 1)   # :127
    out = F.conv2d(out, p["target_conv.weight"], p["target_conv.bias"], padding=2)                     # :130
    out = out.mean(dim=(2, 3)).reshape(B, -1)                                                          # :131-132
    out = F.linear(out, p["readout_dense.weight"], p["readout_dense.bias"])                            # :133
    if testmode:
        return out, torch.stack(states, 1), torch.stack(gates, 1)
    return out
