"""InT forward written with the SAME PyTorch library operators the reference issues.  TEST / BASELINE
INFRASTRUCTURE ONLY (same rules as int_oracle.py: only tests/ and bench.py's reference legs import it).

`int_oracle.forward` states the arithmetic as tensor algebra (einsum gates, hand-written batch
statistics) so that every intermediate can be checkpointed.  This file states the same circuit
through the operator sequence of the reference's eager path, so that timing it on a device is a
fair stand-in for "the reference module on that device" (the reference GPU eager bar):

  models/InT.py:212-213   Conv3d(3 -> C, 1x1x1) over the whole clip + nl          -> F.conv3d, F.softplus
  models/InT.py:148       a_w_gate / a_u_gate: nn.Conv2d(C, C, 1) with bias       -> F.conv2d 1x1
  models/InT.py:161,172   F.conv2d(., w_inh / w_exc, padding) + nn.BatchNorm2d(eps=1e-3,
                          track_running_stats=False)                              -> F.conv2d, F.batch_norm(training=True)
  models/InT.py:162-175   point-wise gating: one ATen kernel per operator, as autograd records them
  models/InT.py:236-241   readout_conv, cat with the goal map, target_conv 5x5, avg_pool2d, Linear

It is device-agnostic (runs wherever its tensors live) and is pinned by tests/test_oracle_golden.py
against int_oracle.forward (which is itself pinned against the reference's fp64 outputs in tests/golden/).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .int_oracle import BN_EPS


def _c1(W):
    return W.reshape(W.shape[0], W.shape[1], 1, 1)


def _s(v):
    return v.view(-1, 1, 1)  # the reference keeps alpha / mu / kappa / gamma as [C,1,1] parameters


def forward(x, p, variant="int", nl=F.softplus):
    """x [B,3,T,H,W], p = int_oracle.unpack(state_dict) -> logits [B,1].  Autograd records the tape."""
    B, _, T, H, W = x.shape
    C = p["bp"].numel()
    xbn = nl(F.conv3d(x, p["Wp"].reshape(C, 3, 1, 1, 1), p["bp"]))                      # InT.py:212-213
    E = torch.zeros((B, C, H, W), dtype=x.dtype, device=x.device)                          # :217-218
    I = torch.zeros((B, C, H, W), dtype=x.dtype, device=x.device)
    pad = p["w_exc"].shape[-1] // 2
    alpha, mu, kappa, gamma = _s(p["alpha"]), _s(p["mu"]), _s(p["kappa"]), _s(p["gamma"])
    for t in range(T):
        X = xbn[:, :, t]
        A = torch.sigmoid(F.conv2d(X, _c1(p["Wa_w"]), p["ba_w"]) + F.conv2d(E, _c1(p["Wa_u"]), p["ba_u"]))  # :148
        Ge = A * E                                                                          # :153
        S = I if variant == "int" else A                                                    # :157 / ffhgru_hierarchy.py:147
        inh = F.batch_norm(F.conv2d(Ge, p["w_inh"], padding=pad), None, None, p["bn0_w"], p["bn0_b"],
                           training=True, eps=BN_EPS)                                       # :161
        Ihat = nl(X - nl(inh * (alpha * S + mu)))                                           # :162
        Gi = torch.sigmoid(F.conv2d(X, _c1(p["Wi_w"]), p["bi_w"]) + F.conv2d(S, _c1(p["Wi_u"]), p["bi_u"]))  # :165
        I = (1 - Gi) * I + Gi * Ihat                                                        # :166
        Gx = torch.sigmoid(F.conv2d(S, _c1(p["We_w"]), p["be_w"]) + F.conv2d(Ge, _c1(p["We_u"]), p["be_u"]))  # :171
        exc = F.batch_norm(F.conv2d(I, p["w_exc"], padding=pad), None, None, p["bn1_w"], p["bn1_b"],
                           training=True, eps=BN_EPS)                                       # :172
        Ehat = nl(exc * (kappa * I + gamma))                                                # :173
        E = (1 - Gx) * E + Gx * Ehat                                                        # :175
    out = torch.cat([F.conv2d(E, p["Wr"].reshape(1, C, 1, 1), p["br"]), x[:, 2, 0][:, None]], 1)  # :236
    out = F.conv2d(out, p["Wt"], p["bt"], padding=p["Wt"].shape[-1] // 2)                   # :238
    out = F.avg_pool2d(out, kernel_size=out.shape[2:]).reshape(B, -1)                       # :239-240
    return F.linear(out, p["Wd"], p["bd"])                                                  # :241
