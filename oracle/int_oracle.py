"""CPU oracle for the InT / hGRU recurrent hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the arithmetic that the reference
performs in

  models/InT.py:145-179            rCell.forward           (cell step, eqs 1-13)
  models/InT.py:210-245            InT.forward             (stem, time loop, readout)
  models/ffhgru_hierarchy.py:135-173  hConvGRUCell.forward (S = att_gate variant)
  models/ffhgru_hierarchy.py:211-276  FFhGRU.forward

written as plain tensor algebra on the CPU (torch CPU tensors, any float dtype;
float64 is the arbiter).  It has two halves:

  * `forward(...)`          the time loop, returning logits and every per-step
                            tensor the CUDA path checkpoints (I_t, E_t, Y0_t, Y1_t,
                            BN statistics) so a parity failure can be localised.
  * `backward_manual(...)`  the hand-derived reverse-time loop (eqs b1-b12 of
                            SURVEY.md section 8) -- exactly the decomposition the
                            CUDA kernels implement -- with no autograd.

Parity pinning: the reference repository holds no golden vectors or tests for
this path, so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
`oracle/gen_golden.py` imports the reference modules from /root/reference,
runs them in fp64, and commits logits + all parameter gradients under
`tests/golden/`.  `tests/test_oracle_golden.py` checks this file against those
fixtures (and `backward_manual` against autograd through `forward`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module.  The product path
(pathtracker_models_b200) never does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # models/InT.py:102  nn.BatchNorm2d(eps=1e-03, track_running_stats=False)

# canonical (reference state_dict) parameter names used by the oracle
GATES = ("a_w", "a_u", "i_w", "i_u", "e_w", "e_u")


def act_fn(kind: str):
    """activation + derivative; softplus is torch's beta=1, threshold=20 (InT.py:145,208)."""
    if kind == "softplus":
        def f(x):
            return torch.where(x > 20, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))

        def df(x):
            return torch.where(x > 20, torch.ones_like(x), torch.sigmoid(x))
        return f, df
    if kind == "tanh":  # engine.py:138-146 InT_tanh
        return torch.tanh, (lambda x: 1 - torch.tanh(x) ** 2)
    raise ValueError(kind)


def unpack(sd: dict, dtype=torch.float64) -> dict:
    """reference state_dict -> flat oracle parameter dict (2-D gate matrices, 1-D scalars)."""
    g = lambda k: sd[k].detach().to(dtype).clone()
    p = {
        "Wp": g("preproc.weight").reshape(-1, 3), "bp": g("preproc.bias"),
        "w_exc": g("unit1.w_exc"),
        "gamma": g("unit1.gamma").flatten(), "kappa": g("unit1.kappa").flatten(),
        "bn0_w": g("unit1.bn.0.weight"), "bn0_b": g("unit1.bn.0.bias"),
        "bn1_w": g("unit1.bn.1.weight"), "bn1_b": g("unit1.bn.1.bias"),
        "Wr": g("readout_conv.weight").reshape(1, -1), "br": g("readout_conv.bias"),
        "Wt": g("target_conv.weight"), "bt": g("target_conv.bias"),
        "Wd": g("readout_dense.weight"), "bd": g("readout_dense.bias"),
    }
    if "unit1.w_inh" in sd:
        p["w_inh"] = g("unit1.w_inh")
    p["alpha"] = g("unit1.alpha").flatten()
    p["mu"] = g("unit1.mu").flatten()
    for n in GATES:
        W = g(f"unit1.{n}_gate.weight")
        p[f"W{n}"] = W.reshape(W.shape[0], W.shape[1])
        p[f"b{n}"] = g(f"unit1.{n}_gate.bias")
    return p


def mv(W, x):
    """1x1 conv without bias: per-pixel C_out x C_in mat-vec on [B,C,H,W]."""
    return torch.einsum("oi,bihw->bohw", W, x)


def mvT(W, g):
    return torch.einsum("oi,bohw->bihw", W, g)


def outer(g, x):
    """weight gradient of a 1x1 conv: sum over (b,h,w) of g (x) x."""
    return torch.einsum("bohw,bihw->oi", g, x)


def csum(x):
    return x.sum(dim=(0, 2, 3))


def cv(v):
    return v.view(1, -1, 1, 1)


def bn_stats(Y):
    """batch statistics over (b,y,x), biased variance (InT.py:102 track_running_stats=False)."""
    mean = Y.mean(dim=(0, 2, 3))
    var = ((Y - cv(mean)) ** 2).mean(dim=(0, 2, 3))
    rstd = 1.0 / torch.sqrt(var + BN_EPS)
    return mean, rstd


def stem(x_t, p, act):
    """InT.py:212-213: Conv3d 1x1x1 (3->C) + nl, one frame.  x_t [B,3,H,W]."""
    pre = mv(p["Wp"], x_t) + cv(p["bp"])
    return act(pre), pre


def cell_forward(X, I, E, p, variant="int", act_kind="softplus", no_inh=False):
    """One cell step, eqs (1)-(13).  Returns dict of every intermediate."""
    act, _ = act_fn(act_kind)
    r = {}
    A = torch.sigmoid(mv(p["Wa_w"], X) + cv(p["ba_w"]) + mv(p["Wa_u"], E) + cv(p["ba_u"]))  # InT.py:148
    Ge = A * E                                                                              # :153
    pad = p["w_exc"].shape[-1] // 2
    if not no_inh:
        S = I if variant == "int" else A      # InT.py:157  vs  ffhgru_hierarchy.py:147
        Y0 = F.conv2d(Ge, p["w_inh"], padding=pad)                                          # :161
        m0, r0 = bn_stats(Y0)
        xh0 = (Y0 - cv(m0)) * cv(r0)
        U = cv(p["bn0_w"]) * xh0 + cv(p["bn0_b"])
        P = U * (cv(p["alpha"]) * S + cv(p["mu"]))                                          # :162
        Q = act(P)
        Ihat = act(X - Q)
        Gi = torch.sigmoid(mv(p["Wi_w"], X) + cv(p["bi_w"]) + mv(p["Wi_u"], S) + cv(p["bi_u"]))  # :165
        I2 = (1 - Gi) * I + Gi * Ihat                                                       # :166
        r.update(Y0=Y0, m0=m0, r0=r0, U=U, P=P, Q=Q, Ihat=Ihat, Gi=Gi)
    else:
        I2, S = Ge, E                                                                       # :168
    Gex = torch.sigmoid(mv(p["We_w"], S) + cv(p["be_w"]) + mv(p["We_u"], Ge) + cv(p["be_u"]))    # :171
    Y1 = F.conv2d(I2, p["w_exc"], padding=pad)                                              # :172
    m1, r1 = bn_stats(Y1)
    V = cv(p["bn1_w"]) * ((Y1 - cv(m1)) * cv(r1)) + cv(p["bn1_b"])
    R = V * (cv(p["kappa"]) * I2 + cv(p["gamma"]))                                          # :173
    Ehat = act(R)
    E2 = (1 - Gex) * E + Gex * Ehat                                                         # :175
    r.update(A=A, Ge=Ge, S=S, Gex=Gex, Y1=Y1, m1=m1, r1=r1, V=V, R=R, Ehat=Ehat, I2=I2, E2=E2)
    return r


def readout(E, goal, p):
    """InT.py:236-241.  goal = x[:, 2, 0]  [B,H,W]."""
    r = mv(p["Wr"], E) + cv(p["br"])
    z = torch.cat([r, goal[:, None]], 1)
    o = F.conv2d(z, p["Wt"], p["bt"], padding=p["Wt"].shape[-1] // 2)
    o = o.mean(dim=(2, 3)).reshape(E.shape[0], -1)
    return o @ p["Wd"].t() + p["bd"]


def forward(x, p, variant="int", act_kind="softplus", no_inh=False, keep=True, testmode=False):
    """InT.forward / FFhGRU.forward.  x [B,3,T,H,W].  Returns (logits, saved)."""
    act, _ = act_fn(act_kind)
    B, _, T, H, W = x.shape
    C = p["bp"].numel()
    I = x.new_zeros(B, C, H, W)
    E = x.new_zeros(B, C, H, W)
    saved = {"I": [I], "E": [E], "Y0": [], "Y1": [], "m0": [], "r0": [], "m1": [], "r1": [],
             "gates": [], "states": []}
    for t in range(T):
        X, _ = stem(x[:, :, t], p, act)
        r = cell_forward(X, I, E, p, variant, act_kind, no_inh)
        I, E = r["I2"], r["E2"]
        if keep:
            saved["I"].append(I); saved["E"].append(E); saved["Y1"].append(r["Y1"])
            saved["m1"].append(r["m1"]); saved["r1"].append(r["r1"])
            if not no_inh:
                saved["Y0"].append(r["Y0"]); saved["m0"].append(r["m0"]); saved["r0"].append(r["r0"])
        if testmode:
            saved["gates"].append(r["A"])
            saved["states"].append(mv(p["Wr"], E) + cv(p["br"]))
    logits = readout(E, x[:, 2, 0], p)
    if not keep:
        saved["I"], saved["E"] = [I], [E]
    return logits, saved


def readout_backward(E, goal, p, dlogits):
    """reverse of `readout`: returns dE and grads of the 6 head parameters."""
    B, C, H, W = E.shape
    g = {}
    r = mv(p["Wr"], E) + cv(p["br"])
    z = torch.cat([r, goal[:, None]], 1)
    ks = p["Wt"].shape[-1]
    o = F.conv2d(z, p["Wt"], p["bt"], padding=ks // 2)
    pooled = o.mean(dim=(2, 3)).reshape(B, -1)              # [B,1]
    g["Wd"] = dlogits.t() @ pooled
    g["bd"] = dlogits.sum(0)
    dpooled = dlogits @ p["Wd"]                             # [B,1]
    do = (dpooled / (H * W)).view(B, 1, 1, 1).expand(B, 1, H, W)
    g["bt"] = do.sum(dim=(0, 2, 3))
    zp = F.pad(z, (ks // 2,) * 4)
    g["Wt"] = torch.stack([torch.stack([(do[:, 0] * zp[:, c, ky:ky + H, kx:kx + W]).sum()
                                        for ky in range(ks) for kx in range(ks)]).view(ks, ks)
                           for c in range(2)])[None]
    dz = F.conv_transpose2d(do.contiguous(), p["Wt"], padding=ks // 2)
    dr = dz[:, :1]
    g["Wr"] = outer(dr, E)
    g["br"] = csum(dr)
    dE = mvT(p["Wr"], dr)
    return dE, g


def backward_manual(x, p, saved, dlogits, variant="int", act_kind="softplus"):
    """Reverse-time loop, eqs (b1)-(b12) of SURVEY.md section 8.  No autograd.

    Uses only what the CUDA path keeps: I_t, E_t (t = 0..T), Y0_t, Y1_t and the BN
    statistics; everything else is recomputed pointwise.  Returns dict of
    gradients keyed like `unpack`.
    """
    act, dact = act_fn(act_kind)
    B, _, T, H, W = x.shape
    N = B * H * W
    C = p["bp"].numel()
    pad = p["w_exc"].shape[-1] // 2
    g = {k: torch.zeros_like(v) for k, v in p.items()}
    dE, gh = readout_backward(saved["E"][T], x[:, 2, 0], p, dlogits)
    g.update(gh)
    dI = torch.zeros_like(dE)
    for t in range(T - 1, -1, -1):
        I, E, I2 = saved["I"][t], saved["E"][t], saved["I"][t + 1]
        Y0, Y1 = saved["Y0"][t], saved["Y1"][t]
        m0, r0, m1, r1 = saved["m0"][t], saved["r0"][t], saved["m1"][t], saved["r1"][t]
        # ---- pointwise recompute of the forward step
        X, pre = stem(x[:, :, t], p, act)
        A = torch.sigmoid(mv(p["Wa_w"], X) + cv(p["ba_w"]) + mv(p["Wa_u"], E) + cv(p["ba_u"]))
        Ge = A * E
        S = I if variant == "int" else A
        xh0 = (Y0 - cv(m0)) * cv(r0)
        U = cv(p["bn0_w"]) * xh0 + cv(p["bn0_b"])
        aSm = cv(p["alpha"]) * S + cv(p["mu"])
        P = U * aSm
        Q = act(P)
        Ihat = act(X - Q)
        Gi = torch.sigmoid(mv(p["Wi_w"], X) + cv(p["bi_w"]) + mv(p["Wi_u"], S) + cv(p["bi_u"]))
        Gex = torch.sigmoid(mv(p["We_w"], S) + cv(p["be_w"]) + mv(p["We_u"], Ge) + cv(p["be_u"]))
        xh1 = (Y1 - cv(m1)) * cv(r1)
        V = cv(p["bn1_w"]) * xh1 + cv(p["bn1_b"])
        kIg = cv(p["kappa"]) * I2 + cv(p["gamma"])
        R = V * kIg
        Ehat = act(R)
        dEn, dIn = dE, dI                       # incoming dE', dI'
        # ---- (b1)-(b2)
        dGex = dEn * (Ehat - E)
        dE = dEn * (1 - Gex)
        dR = dEn * Gex * dact(R)
        dV = dR * kIg
        dIn = dIn + dR * V * cv(p["kappa"])
        g["kappa"] += csum(dR * V * I2)
        g["gamma"] += csum(dR * V)
        # ---- (b3) BN1 backward
        sdV, sdVx = csum(dV), csum(dV * xh1)
        g["bn1_w"] += sdVx
        g["bn1_b"] += sdV
        dY1 = cv(p["bn1_w"] * r1) * (dV - cv(sdV / N) - xh1 * cv(sdVx / N))
        # ---- (b4) dgrad / wgrad of the excitation conv
        dIn = dIn + F.conv_transpose2d(dY1, p["w_exc"], padding=pad)
        g["w_exc"] += torch.nn.grad.conv2d_weight(I2, p["w_exc"].shape, dY1, padding=pad)
        # ---- (b5) exc gate
        ge = dGex * Gex * (1 - Gex)
        dS = mvT(p["We_w"], ge)
        dGe = mvT(p["We_u"], ge)
        g["We_w"] += outer(ge, S); g["We_u"] += outer(ge, Ge)
        g["be_w"] += csum(ge); g["be_u"] += csum(ge)
        # ---- (b6) inhibition blend
        dGi = dIn * (Ihat - I)
        dI = dIn * (1 - Gi)
        z = dIn * Gi * dact(X - Q)
        dX = z
        # ---- (b7) inh gate
        gi = dGi * Gi * (1 - Gi)
        dX = dX + mvT(p["Wi_w"], gi)
        dS = dS + mvT(p["Wi_u"], gi)
        g["Wi_w"] += outer(gi, X); g["Wi_u"] += outer(gi, S)
        g["bi_w"] += csum(gi); g["bi_u"] += csum(gi)
        # ---- (b8)
        dP = -z * dact(P)
        dU = dP * aSm
        dS = dS + dP * U * cv(p["alpha"])
        g["alpha"] += csum(dP * U * S)
        g["mu"] += csum(dP * U)
        # ---- (b9) BN0 backward, dgrad / wgrad of the inhibition conv
        sdU, sdUx = csum(dU), csum(dU * xh0)
        g["bn0_w"] += sdUx
        g["bn0_b"] += sdU
        dY0 = cv(p["bn0_w"] * r0) * (dU - cv(sdU / N) - xh0 * cv(sdUx / N))
        dGe = dGe + F.conv_transpose2d(dY0, p["w_inh"], padding=pad)
        g["w_inh"] += torch.nn.grad.conv2d_weight(Ge, p["w_inh"].shape, dY0, padding=pad)
        # ---- (b10)
        dA = dGe * E
        dE = dE + dGe * A
        if variant == "int":
            dI = dI + dS
        else:
            dA = dA + dS
        # ---- (b11) attention gate
        a = dA * A * (1 - A)
        dX = dX + mvT(p["Wa_w"], a)
        dE = dE + mvT(p["Wa_u"], a)
        g["Wa_w"] += outer(a, X); g["Wa_u"] += outer(a, E)
        g["ba_w"] += csum(a); g["ba_u"] += csum(a)
        # ---- (b12) stem
        dpre = dX * dact(pre)
        g["Wp"] += outer(dpre, x[:, :, t])
        g["bp"] += csum(dpre)
    return g


def bce_with_logits_grad(logits, y):
    """d/dlogits of mean BCEWithLogits (mainclean.py:160,190 criterion)."""
    return (torch.sigmoid(logits) - y) / logits.numel()


# map oracle keys -> reference state_dict keys (for comparing gradients)
def grad_to_state_keys(g: dict, sd: dict) -> dict:
    out = {}
    m = {"Wp": "preproc.weight", "bp": "preproc.bias", "w_exc": "unit1.w_exc", "w_inh": "unit1.w_inh",
         "alpha": "unit1.alpha", "mu": "unit1.mu", "gamma": "unit1.gamma", "kappa": "unit1.kappa",
         "bn0_w": "unit1.bn.0.weight", "bn0_b": "unit1.bn.0.bias",
         "bn1_w": "unit1.bn.1.weight", "bn1_b": "unit1.bn.1.bias",
         "Wr": "readout_conv.weight", "br": "readout_conv.bias", "Wt": "target_conv.weight",
         "bt": "target_conv.bias", "Wd": "readout_dense.weight", "bd": "readout_dense.bias"}
    for n in GATES:
        m[f"W{n}"] = f"unit1.{n}_gate.weight"
        m[f"b{n}"] = f"unit1.{n}_gate.bias"
    for k, v in g.items():
        if m[k] in sd:
            out[m[k]] = v.reshape(sd[m[k]].shape)
    return out
