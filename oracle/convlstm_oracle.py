"""CPU oracle for the ConvLSTM baseline (models/convlstm.py).  TEST INFRASTRUCTURE ONLY.

Restates ConvLSTMCell.forward (models/convlstm.py:84-90) and ConvLSTM.forward (:116-173, bptt branch):
Gabor stem conv0 -> square, T cell steps on the time-invariant input, batch-stat BatchNorm + 1x1 conv head,
and the two vector-Jacobian products of the Jacobian penalty (:152-163).  The input-path convolutions
Wx*(x) + b are hoisted out of the time loop (the reference recomputes them every step; identical values).

`rbp_grads` restates grad_method='rbp' (:124-135) with the Neumann series of `dummyhgru.backward` (:9-54).

Pinned against outputs of the reference itself: oracle/gen_golden.py -> tests/golden/convlstm_*.npz.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
GATES = ("i", "f", "c", "o")   # stacking order of the 4 x 25 gate channels used by the CUDA path


def unpack(sd, dtype=torch.float64):
    g = lambda k: sd[k].detach().to(dtype).clone()
    p = {"w0": g("conv0.weight"), "b0": g("conv0.bias"), "bn_w": g("bn.weight"), "bn_b": g("bn.bias"),
         "w6": g("conv6.weight"), "b6": g("conv6.bias")}
    for k in GATES:
        p["wx" + k] = g(f"unit1.Wx{k}.weight")
        p["bx" + k] = g(f"unit1.Wx{k}.bias")
        p["wh" + k] = g(f"unit1.Wh{k}.weight")
    return p


def cell(Xg, h, c, p, pad):
    """Xg: dict of hoisted input terms; returns h', c' and the gate activations."""
    i = torch.sigmoid(Xg["i"] + F.conv2d(h, p["whi"], padding=pad))      # convlstm.py:85
    f = torch.sigmoid(Xg["f"] + F.conv2d(h, p["whf"], padding=pad))      # :86
    g = torch.tanh(Xg["c"] + F.conv2d(h, p["whc"], padding=pad))         # :87
    c2 = f * c + i * g
    o = torch.sigmoid(Xg["o"] + F.conv2d(h, p["who"], padding=pad))      # :88
    h2 = o * torch.tanh(c2)                                              # :89
    return h2, c2, (i, f, g, o)


def forward(x, p, timesteps, keep=False):
    """x [B,1,H,W] -> (output [B,2,H,W], saved)."""
    pad = p["whi"].shape[-1] // 2
    xt = F.conv2d(x, p["w0"], p["b0"], padding=p["w0"].shape[-1] // 2) ** 2      # :118-119
    Xg = {k: F.conv2d(xt, p["wx" + k], p["bx" + k], padding=pad) for k in GATES}
    h = torch.zeros_like(xt)
    c = torch.zeros_like(xt)
    saved = {"h": [h], "c": [c], "gates": [], "xt": xt}
    for _ in range(timesteps):
        h, c, gates = cell(Xg, h, c, p, pad)
        if keep:
            saved["h"].append(h); saved["c"].append(c); saved["gates"].append(gates)
    saved["h_last"], saved["c_last"] = h, c
    mean = h.mean(dim=(0, 2, 3))
    var = ((h - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
    hn = (h - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    hn = hn * p["bn_w"].view(1, -1, 1, 1) + p["bn_b"].view(1, -1, 1, 1)             # :146
    out = F.conv2d(hn, p["w6"], p["b6"])                                             # :147
    return out, saved


def jv_penalty(x, p, timesteps):
    """(1^T dh_T/dh_{T-1} - 0.9)+^2 + (1^T dc_T/dc_{T-1} - 0.9)+^2   (convlstm.py:152-163, no graph kept).

    As in the reference, h_{T-1} and c_{T-1} are the live graph tensors, so the second product also
    contains the path c_{T-1} -> h_{T-1} = o tanh(c_{T-1}) -> gates_T -> c_T, not only f_T."""
    pad = p["whi"].shape[-1] // 2
    xt = F.conv2d(x, p["w0"], p["b0"], padding=p["w0"].shape[-1] // 2) ** 2
    Xg = {k: F.conv2d(xt, p["wx" + k], p["bx" + k], padding=pad) for k in GATES}
    h = torch.zeros_like(xt).requires_grad_(True)      # T == 1 is an error in the reference (needs T >= 2)
    c = torch.zeros_like(xt).requires_grad_(True)
    for _ in range(timesteps - 1):
        h, c, _ = cell(Xg, h, c, p, pad)
    h2, c2, _ = cell(Xg, h, c, p, pad)
    ones = torch.ones_like(h2)
    jh = torch.autograd.grad(h2, h, grad_outputs=ones, retain_graph=True)[0]
    jc = torch.autograd.grad(c2, c, grad_outputs=ones)[0]
    return ((jh - 0.9).clamp(0) ** 2 + (jc - 0.9).clamp(0) ** 2).detach()


def jp_total_grads(x, p, timesteps, target, grad_method="bptt", num_iter=50, weight=10.0):
    """jacobian_penalty=True (models/convlstm.py:152-163 with create_graph=True): gradients of
    loss + weight * penalty.mean() (mainclean.py:192-196) by double back-propagation.  `p` is keyed like the state_dict,
    leaf tensors with requires_grad=True.  Returns (output, loss, penalty, grads).

    bptt (:137-144): h_{T-1}, c_{T-1} are live graph tensors, so the penalty's gradient runs back through every step and
    the second product contains the path c_{T-1} -> h_{T-1} -> gates_T -> c_T.  rbp (:124-135): they are detached leaves,
    the second product is f_T alone and the penalty reaches the parameters through the last step only; the loss
    gradient is the Neumann series of `dummyhgru` (rbp_grads)."""
    u = F.conv2d(x, p["conv0.weight"], p["conv0.bias"], padding=3)
    xt = u * u
    pad = (p["unit1.Whi.weight"].shape[-1] - 1) // 2
    h = torch.zeros_like(xt)
    c = torch.zeros_like(xt)
    if grad_method == "rbp":
        with torch.no_grad():
            for _ in range(timesteps - 1):
                h, c = _cell_sd(xt, h, c, p, pad)
        h2 = h.detach().requires_grad_()
        c2 = c.detach().requires_grad_()
    else:
        for _ in range(timesteps - 1):
            h, c = _cell_sd(xt, h, c, p, pad)
        h2, c2 = h, c
    hT, cT = _cell_sd(xt, h2, c2, p, pad)
    ones = torch.ones_like(hT)
    jh = torch.autograd.grad(hT, h2, grad_outputs=ones, retain_graph=True, create_graph=True)[0]
    jc = torch.autograd.grad(cT, c2, grad_outputs=ones, retain_graph=True, create_graph=True)[0]
    pen = (jh - 0.9).clamp(0) ** 2 + (jc - 0.9).clamp(0) ** 2
    keys = list(p)
    if grad_method == "rbp":
        out, loss, grads, _ = rbp_grads(x, p, timesteps, target, num_iter)
        pg = torch.autograd.grad(pen.mean() * weight, [p[k] for k in keys], allow_unused=True)
        for k, g in zip(keys, pg):
            if g is not None:
                grads[k] = g if grads.get(k) is None else grads[k] + g
        return out, loss, pen.detach(), grads
    out = F.conv2d(F.batch_norm(hT, None, None, p["bn.weight"], p["bn.bias"], True, 0.0, 1e-3), p["conv6.weight"], p["conv6.bias"])
    loss = F.cross_entropy(out, target)
    wrt = [p[k] for k in keys] + ([x] if x.requires_grad else [])          # bptt: also d total / d x when asked for (saliency)
    pg = torch.autograd.grad(loss + pen.mean() * weight, wrt, allow_unused=True)
    grads = {k: g for k, g in zip(keys, pg)}
    if x.requires_grad:
        grads["__x__"] = pg[-1]
    return out.detach(), loss.item(), pen.detach(), grads


STATE_KEYS = {"w0": "conv0.weight", "b0": "conv0.bias", "bn_w": "bn.weight", "bn_b": "bn.bias",
              "w6": "conv6.weight", "b6": "conv6.bias"}
for _k in GATES:
    STATE_KEYS["wx" + _k] = f"unit1.Wx{_k}.weight"
    STATE_KEYS["bx" + _k] = f"unit1.Wx{_k}.bias"
    STATE_KEYS["wh" + _k] = f"unit1.Wh{_k}.weight"


def rbp_grads(x, p, timesteps, target, num_iter, criterion=None):
    """grad_method='rbp' (models/convlstm.py:124-135 + dummyhgru :9-54): returns (output, loss, grads by parameter
    key of `p`, accepted Neumann terms).  `p` values must be leaf tensors with requires_grad=True."""
    crit = criterion or F.cross_entropy
    u = F.conv2d(x, p["conv0.weight"], p["conv0.bias"], padding=3)
    xt = u * u
    pad = (p["unit1.Whi.weight"].shape[-1] - 1) // 2
    h = torch.zeros_like(xt)
    c = torch.zeros_like(xt)
    with torch.no_grad():                                   # :125-128
        for _ in range(timesteps - 1):
            h, c = _cell_sd(xt, h, c, p, pad)
    h2 = h.detach().requires_grad_()                        # state_2nd_last (:130-131)
    c2 = c.detach().requires_grad_()
    hT, _cT = _cell_sd(xt, h2, c2, p, pad)                  # :133, with graph to the parameters and conv0
    # head: bn (batch statistics, eps 1e-3) + conv6 on a stand-in leaf, to obtain dL/dh_T
    hT_leaf = hT.detach().requires_grad_()
    out = F.conv2d(F.batch_norm(hT_leaf, None, None, p["bn.weight"], p["bn.bias"], True, 0.0, 1e-3), p["conv6.weight"], p["conv6.bias"])
    loss = crit(out, target)
    head_params = [p[k] for k in ("bn.weight", "bn.bias", "conv6.weight", "conv6.bias")]
    gr = torch.autograd.grad(loss, [hT_leaf] + head_params)
    grad = gr[0]
    # dummyhgru.backward (:17-54)
    g_prev, v_prev = grad.clone(), grad.clone()
    normsv = [torch.norm(g_prev).item()]
    terms = 0
    for _ in range(num_iter):
        v = torch.autograd.grad(hT, h2, grad_outputs=v_prev, retain_graph=True)[0]
        normv = torch.norm(v)
        g = g_prev + v
        normg = torch.norm(g)
        if normg > 1 or normv > normsv[-1] or normv < 1e-9:
            break
        v_prev, g_prev = v, g
        normsv.append(normv.item())
        terms += 1
    cell_keys = [k for k in p if k not in ("bn.weight", "bn.bias", "conv6.weight", "conv6.bias")]
    cg = torch.autograd.grad(hT, [p[k] for k in cell_keys], grad_outputs=g_prev, allow_unused=True)
    grads = {k: v for k, v in zip(cell_keys, cg)}
    grads.update({k: v for k, v in zip(("bn.weight", "bn.bias", "conv6.weight", "conv6.bias"), gr[1:])})
    return out.detach(), loss.item(), grads, terms


def _cell_sd(xt, h, c, p, pad):
    """ConvLSTMCell.forward (models/convlstm.py:84-90) on state_dict-keyed parameters."""
    def cx(k):
        return F.conv2d(xt, p[f"unit1.Wx{k}.weight"], p[f"unit1.Wx{k}.bias"], padding=pad)

    def ch(k):
        return F.conv2d(h, p[f"unit1.Wh{k}.weight"], None, padding=pad)
    i = torch.sigmoid(cx("i") + ch("i"))
    f = torch.sigmoid(cx("f") + ch("f"))
    cc = f * c + i * torch.tanh(cx("c") + ch("c"))
    o = torch.sigmoid(cx("o") + ch("o"))
    return o * torch.tanh(cc), cc
