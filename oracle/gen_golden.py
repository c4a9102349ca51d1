"""Generate golden fixtures by RUNNING THE REFERENCE (test infrastructure only).

Runs in the build container only (needs /root/reference, which does not exist on
the GPU box).  Imports the unmodified reference modules with the two shims of
SURVEY.md section 0 (missing super() globals; `.cuda()` on a CPU host), runs them in
float64 (arbiter) and float32 (the reference's own noise floor), and writes
logits + every parameter gradient + the state_dict + the input recipe to
tests/golden/<case>.npz.

    python oracle/gen_golden.py            # regenerates every fixture

Input recipe (reproducible without torch RNG): numpy RandomState(seed).rand(shape)
for x, (rand(B,1) > .5) for y.  Weights: reference constructor under
torch.manual_seed(wseed), then every parameter perturbed by a seeded numpy normal
so that no term of the cell is at a degenerate init value (mu = gamma = 0, ...).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    sys.path.insert(0, REF)
    from models import InT as RInT
    from models import ffhgru_hierarchy as RH
    RInT.FFhGRU = RInT.InT          # super(FFhGRU, self) in InT.__init__
    RInT.hConvGRUCell = RInT.rCell  #                 super(hConvGRUCell, self) in rCell.__init__
    torch.Tensor.cuda = lambda self, *a, **k: self  # SURVEY section 0.2 (CPU host only)
    return RInT, RH


def make_inputs(seed, B, T, H, W):
    rs = np.random.RandomState(seed)
    x = rs.rand(B, 3, T, H, W)
    y = (rs.rand(B, 1) > 0.5).astype(np.float64)
    return x, y


def perturb(model, seed, scale=0.05):
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if n == "unit1.w":            # uninitialised + unused (InT.py:100,118): pin to a constant
                prm.fill_(1.0)
                continue
            if not torch.isfinite(prm).all() or n in ("unit1.alpha", "unit1.mu") and prm.abs().max() > 1e3:
                prm.fill_(0.5)            # no_inh leaves alpha/mu uninitialised (InT.py:112-114)
            prm.add_(torch.from_numpy(rs.normal(0, scale, tuple(prm.shape))).to(prm.dtype))


def run_case(name, ctor, B, T, H, W, seed, wseed, testmode=False):
    x_np, y_np = make_inputs(seed, B, T, H, W)
    res = {"x_seed": seed, "shape": np.array([B, T, H, W])}
    # weights: fp32-representable, so the fixture stores them exactly as float32
    torch.set_default_dtype(torch.float32)
    torch.manual_seed(wseed)
    m32 = ctor()
    perturb(m32, wseed + 1)
    sd32 = {k: v.detach().clone() for k, v in m32.state_dict().items()}
    for k, v in sd32.items():
        res["sd/" + k] = v.numpy()
    grads = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(dt)   # states are created with the default dtype (InT.py:217-218)
        m = ctor().to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in sd32.items()})
        x = torch.from_numpy(x_np).to(dt)
        y = torch.from_numpy(y_np).to(dt)
        out, _ = m(x)
        loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        grads[tag] = {n: prm.grad.detach().double() for n, prm in m.named_parameters() if prm.grad is not None}
        grads[tag]["__logits__"] = out.detach().double()
        if tag == "f64":
            res["logits"] = out.detach().numpy()
            res["loss"] = np.array(loss.item())
            for n, gr in grads[tag].items():
                if n == "__logits__":
                    continue
                big = gr.numel() > 4096   # the two 7x7 banks: float32 storage (6e-8 rel) keeps fixtures small
                res["grad/" + n] = gr.numpy().astype(np.float32 if big else np.float64)
            if testmode:
                with torch.no_grad():
                    o2, states, gates = m(x, testmode=True)
                res["tm_logits"] = o2.numpy()
                res["tm_states"] = states.numpy().astype(np.float32)
                res["tm_gates"] = gates.numpy().astype(np.float32)
    # the fp32 reference's own error vs fp64, per tensor: max|g32-g64| / max|g64|  
    for n, g64 in grads["f64"].items():
        e = (grads["f32"][n] - g64).abs().max().item() / max(g64.abs().max().item(), 1e-300)
        res["f32err/" + n] = np.array(e)
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    worst = max((float(v), k) for k, v in res.items() if k.startswith("f32err/"))
    print(name, "logits", res["logits"].ravel()[:4], "loss", res["loss"], "worst fp32-vs-fp64", worst)


def main():
    RInT, RH = load_reference()
    os.makedirs(OUT, exist_ok=True)
    I = lambda **kw: (lambda: RInT.InT(dimensions=32, timesteps=8, kernel_size=7, **kw))
    run_case("int_small", I(), 2, 4, 32, 32, seed=11, wseed=3, testmode=True)
    run_case("int_tile", I(), 3, 3, 32, 32, seed=12, wseed=4)          # batch not a multiple of anything
    run_case("int_cfg1", I(), 4, 32, 32, 32, seed=13, wseed=5)         # BASELINE configs[0] shape
    run_case("int_64px", I(), 1, 3, 64, 64, seed=14, wseed=6)          # cfg-3 frame size, short
    run_case("int_tanh_small", I(nl=torch.tanh), 2, 4, 32, 32, seed=15, wseed=7)
    run_case("int_noinh_small", I(no_inh=True), 2, 4, 32, 32, seed=16, wseed=8)
    run_case("int_lesion_small", I(lesion_alpha=True, lesion_gamma=True), 2, 4, 32, 32, seed=17, wseed=9)
    H = lambda: RH.FFhGRU(dimensions=32, timesteps=8, kernel_size=7)
    run_case("hgru_small", H, 2, 4, 32, 32, seed=21, wseed=10, testmode=True)
    run_case("hgru_tile", H, 3, 3, 32, 32, seed=22, wseed=11)
    # non-normative fingerprint of SURVEY 8c: seed 0, InT(32,32,7), x = torch.rand right after ctor
    torch.manual_seed(0)
    m = RInT.InT(dimensions=32, timesteps=32, kernel_size=7)
    x = torch.rand(4, 3, 32, 32, 32)
    with torch.no_grad():
        out, _ = m(x)
    np.savez_compressed(os.path.join(OUT, "int_fingerprint.npz"), logits=out.numpy(),
                        init_sums=np.array([v.double().sum().item() for k, v in m.state_dict().items()
                                            if k != "unit1.w"]),
                        init_keys=np.array([k for k in m.state_dict().keys() if k != "unit1.w"]))
    print("fingerprint", out.flatten())
    torch.manual_seed(0)
    h = RH.FFhGRU(dimensions=32, timesteps=32, kernel_size=7)
    np.savez_compressed(os.path.join(OUT, "hgru_fingerprint.npz"),
                        init_sums=np.array([v.double().sum().item() for k, v in h.state_dict().items()]),
                        init_keys=np.array(list(h.state_dict().keys())))


def convlstm_weights(sd_like, seed):
    """deterministic, fp32-representable weights for the ConvLSTM fixtures (numpy recipe, so that the
    large filt_size=15 case does not have to store 1.1 M parameters)."""
    rs = np.random.RandomState(seed)
    out = {}
    for k, v in sd_like.items():
        shape = tuple(v.shape)
        if k.endswith("Wxi.weight") or "unit1.W" in k and k.endswith("weight"):
            fan = shape[1] * shape[2] * shape[3]
            a = rs.normal(0, 1.0 / np.sqrt(fan), shape)
        elif k == "conv0.weight":
            a = v.numpy() + rs.normal(0, 0.02, shape)          # the Gabor bank, perturbed
        elif k == "bn.weight":
            a = 1.0 + rs.normal(0, 0.1, shape)
        elif k == "conv6.weight":
            a = rs.normal(0, 0.3, shape)
        elif k == "conv6.bias":
            a = rs.normal(0, 0.5, shape)
        else:
            a = rs.normal(0, 0.1, shape)
        out[k] = torch.from_numpy(a.astype(np.float32))
    return out


def run_convlstm_case(name, filt, B, T, H, W, seed, store_full, grad_method='bptt', num_iter=50, wh_scale=1.0, jacobian_penalty=False):
    """jacobian_penalty=True: the penalty is built with create_graph=True (models/convlstm.py:157-160) and the training
    loop back-propagates loss + 10 * penalty.mean() (mainclean.py:192-196); the stored gradients are those of that sum."""
    cwd = os.getcwd()
    os.chdir(REF)                                   # ConvLSTM.__init__ loads 'utils/gabor_serre.npy' (convlstm.py:105)
    from models import convlstm as RC
    rs = np.random.RandomState(seed)
    x_np = rs.rand(B, 1, H, W)
    tgt_np = (rs.rand(B, H, W) > 0.5).astype(np.int64)
    res = {"x_seed": seed, "shape": np.array([B, T, H, W]), "filt": np.array(filt)}
    torch.set_default_dtype(torch.float32)
    m0 = RC.ConvLSTM(timesteps=T, filt_size=filt)
    sd32 = convlstm_weights(m0.state_dict(), seed + 1)
    if wh_scale != 1.0:     # rbp fixtures: scale the recurrent weights so that the Neumann series runs several terms
        for k in sd32:
            if ".Wh" in k:
                sd32[k] = sd32[k] * wh_scale
        res["wh_scale"] = np.array(wh_scale)
    res["num_iter"] = np.array(num_iter)
    grads = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(dt)
        m = RC.ConvLSTM(timesteps=T, filt_size=filt, grad_method=grad_method, num_iter=num_iter,
                        jacobian_penalty=jacobian_penalty).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in sd32.items()})
        m.train()
        x = torch.from_numpy(x_np).to(dt)
        tgt = torch.from_numpy(tgt_np)
        out, pen, loss = m(x, 0, 0, tgt, torch.nn.CrossEntropyLoss())
        calls = [0]
        real_grad = torch.autograd.grad
        def counting_grad(*a, **k):
            calls[0] += 1
            return real_grad(*a, **k)
        torch.autograd.grad = counting_grad      # dummyhgru.backward calls it once per Neumann term it tries
        total = loss + pen.mean() * 1e1 if jacobian_penalty else loss      # mainclean.py:192-196
        try:
            total.backward()
        finally:
            torch.autograd.grad = real_grad
        if tag == "f64":
            res["rbp_grad_calls"] = np.array(calls[0])
        grads[tag] = {n: prm.grad.detach().double() for n, prm in m.named_parameters()}
        grads[tag]["__output__"] = out.detach().double()
        grads[tag]["__penalty__"] = pen.detach().double()
        if tag == "f64":
            res["output"] = out.detach().numpy()
            res["loss"] = np.array(loss.item())
            res["penalty"] = pen.detach().numpy()
            prs = np.random.RandomState(seed + 2)
            for n, gr in grads[tag].items():
                if n.startswith("__"):
                    continue
                if store_full:
                    res["grad/" + n] = gr.numpy().astype(np.float32 if gr.numel() > 4096 else np.float64)
                else:   # compact pin: 4 seeded random projections + max-abs
                    v = prs.normal(0, 1, (4, gr.numel()))
                    res["proj/" + n] = v @ gr.numpy().ravel()
                    res["gmax/" + n] = np.array(gr.abs().max().item())
    for n, g64 in grads["f64"].items():
        e = (grads["f32"][n] - g64).abs().max().item() / max(g64.abs().max().item(), 1e-300)
        res["f32err/" + n] = np.array(e)
    if store_full:
        for k, v in sd32.items():
            res["sd/" + k] = v.numpy()
    res["wseed"] = np.array(seed + 1)
    res["gabor"] = m0.state_dict()["conv0.weight"].numpy()       # the reference's Gabor bank (data fixture, 4.9 KB)
    torch.set_default_dtype(torch.float32)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    worst = max((float(v), k) for k, v in res.items() if k.startswith("f32err/"))
    print(name, "loss", res["loss"], "penalty max", res["penalty"].max(), "worst fp32-vs-fp64", worst)


def run_gru_case(name, B, T, H, W, seed):
    """kys.GRU (utils/engine.py:147-153: dimensions = 64, kernel 7).  Weights are the reference's own default init under
    torch.manual_seed(seed): the mirror module must reproduce them bit for bit, so only the seed is stored; gradients
    of the three 64x128x7x7 banks are pinned through seeded random projections."""
    from models import kys
    rs = np.random.RandomState(seed)
    x_np = rs.rand(B, 3, T, H, W)
    y_np = (rs.rand(B, 1) > 0.5).astype(np.float64)
    res = {"x_seed": seed, "shape": np.array([B, T, H, W])}
    grads = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        torch.set_default_dtype(torch.float32)
        torch.manual_seed(seed)
        m = kys.GRU(dimensions=64, timesteps=T, kernel_size=7)
        if tag == "f64":
            res["init_sums"] = np.array([v.double().sum().item() for v in m.state_dict().values()])
            res["init_keys"] = np.array(list(m.state_dict().keys()))
        m = m.to(dt)
        x = torch.from_numpy(x_np).to(dt)
        y = torch.from_numpy(y_np).to(dt)
        out, pen = m(x)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        grads[tag] = {n: prm.grad.detach().double() for n, prm in m.named_parameters() if prm.grad is not None}
        grads[tag]["__logits__"] = out.detach().double()
        if tag == "f64":
            res["logits"] = out.detach().numpy()
            res["loss"] = np.array(loss.item())
            with torch.no_grad():
                o2, states, gates = m(x, testmode=True)
            res["tm_states"] = states.numpy().astype(np.float32)
            res["tm_gates_sum"] = np.array(gates.double().sum().item())
            res["tm_gates_last"] = gates[:, -1, :4].numpy().astype(np.float32)
            prs = np.random.RandomState(seed + 2)
            for n, gr in grads[tag].items():
                if n.startswith("__"):
                    continue
                if gr.numel() <= 4096:
                    res["grad/" + n] = gr.numpy()
                else:
                    v = prs.normal(0, 1, (4, gr.numel()))
                    res["proj/" + n] = v @ gr.numpy().ravel()
                    res["gmax/" + n] = np.array(gr.abs().max().item())
    for n, g64 in grads["f64"].items():
        e = (grads["f32"][n] - g64).abs().max().item() / max(g64.abs().max().item(), 1e-300)
        res["f32err/" + n] = np.array(e)
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    worst = max((float(v), k) for k, v in res.items() if k.startswith("f32err/"))
    print(name, "loss", res["loss"], "worst fp32-vs-fp64", worst)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gru":
        load_reference()
        torch.Tensor.cuda = lambda s, *a, **k: s
        run_gru_case("gru_small", 2, 3, 32, 32, seed=51)
        run_gru_case("gru_b3", 3, 2, 16, 64, seed=52)
    elif len(sys.argv) > 1 and sys.argv[1] == "convlstm":
        load_reference()
        run_convlstm_case("convlstm_k7_small", 7, 2, 3, 32, 32, seed=31, store_full=True)
        run_convlstm_case("convlstm_k15_small", 15, 2, 3, 32, 32, seed=32, store_full=False)
        run_convlstm_case("convlstm_k15_b3", 15, 3, 4, 32, 32, seed=33, store_full=False)
    elif len(sys.argv) > 1 and sys.argv[1] == "convlstm_rbp":
        load_reference()
        run_convlstm_case("convlstm_k7_rbp", 7, 2, 4, 32, 32, seed=41, store_full=True, grad_method='rbp', num_iter=50)
        run_convlstm_case("convlstm_k7_rbp_trunc", 7, 2, 3, 32, 32, seed=42, store_full=True, grad_method='rbp', num_iter=2)
    elif len(sys.argv) > 1 and sys.argv[1] == "convlstm_jp":
        # jacobian_penalty=True: second-order graph through the last cell step, both gradient methods
        load_reference()
        run_convlstm_case("convlstm_k7_jp", 7, 2, 3, 32, 32, seed=71, store_full=True, jacobian_penalty=True)
        run_convlstm_case("convlstm_k7_jp_t2", 7, 2, 2, 16, 32, seed=72, store_full=True, jacobian_penalty=True)
        run_convlstm_case("convlstm_k15_jp", 15, 2, 4, 32, 32, seed=73, store_full=False, jacobian_penalty=True)
        run_convlstm_case("convlstm_k7_rbp_jp", 7, 2, 4, 32, 32, seed=74, store_full=True, grad_method='rbp', num_iter=50,
                          jacobian_penalty=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "long":
        # the horizons the benchmark is quoted on (BASELINE.json configs[1]-[4]): T = 64 at 32x32, T = 128 at 64x64,
        # ConvLSTM timesteps = 64 (models/convlstm.py:138), kys.GRU T = 64.  Minutes of CPU time each.
        which = sys.argv[2:] or ["int", "hgru", "int128", "convlstm", "gru"]
        RInT, RH = load_reference()
        if "int" in which:
            run_case("int_t64", lambda: RInT.InT(dimensions=32, timesteps=64, kernel_size=7), 8, 64, 32, 32, seed=61, wseed=23)
        if "hgru" in which:
            run_case("hgru_t64", lambda: RH.FFhGRU(dimensions=32, timesteps=64, kernel_size=7), 8, 64, 32, 32, seed=62, wseed=24)
        if "int128" in which:
            run_case("int_t128_64px", lambda: RInT.InT(dimensions=32, timesteps=128, kernel_size=7), 2, 128, 64, 64,
                     seed=63, wseed=25)
        if "convlstm" in which:
            run_convlstm_case("convlstm_k15_t64", 15, 2, 64, 32, 32, seed=64, store_full=False)
        if "gru" in which:
            run_gru_case("gru_t64", 2, 64, 32, 32, seed=65)
    elif len(sys.argv) > 1 and sys.argv[1] == "convlstm_rbp_long":
        # recurrent back-propagation at configs[4]'s horizon (timesteps = 64, filt 15): 63 steps under no_grad, Neumann series on the last
        load_reference()
        run_convlstm_case("convlstm_k15_rbp_t64", 15, 2, 64, 32, 32, seed=75, store_full=False, grad_method='rbp', num_iter=15)
    elif len(sys.argv) > 1 and sys.argv[1] == "convlstm_jp_long":
        # the in-graph Jacobian penalty at configs[4]'s horizon: its second-order seeds travel back through all 64 steps
        load_reference()
        run_convlstm_case("convlstm_k15_jp_t64", 15, 2, 64, 32, 32, seed=76, store_full=False, jacobian_penalty=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "long_variants":
        # the InT variants utils/engine.py builds (nl = tanh :138-146, no_inh :86-93, lesions :94-137) at the headline horizon T = 64
        RInT, RH = load_reference()
        run_case("int_tanh_t64", lambda: RInT.InT(dimensions=32, timesteps=64, kernel_size=7, nl=torch.tanh), 4, 64, 32, 32, seed=66, wseed=26)
        run_case("int_noinh_t64", lambda: RInT.InT(dimensions=32, timesteps=64, kernel_size=7, no_inh=True), 4, 64, 32, 32, seed=67, wseed=27)
        run_case("int_lesion_t64", lambda: RInT.InT(dimensions=32, timesteps=64, kernel_size=7, lesion_mu=True, lesion_kappa=True),
                 4, 64, 32, 32, seed=68, wseed=28)
    else:
        main()
