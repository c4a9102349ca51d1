"""Poor man's memcheck (compute-sanitizer is closed on this pool, profiles/r2a_sanitizer_closed.txt): every device buffer
the Python layer hands to the C ABI -- workspace, checkpoints, outputs, gradients -- is allocated between two canary bands;
after a forward + backward of each model family and precision mode the bands must be untouched.  Catches out-of-bounds
WRITES of up to 1 MiB either side of any caller-owned buffer (the library allocates nothing itself)."""
import contextlib

import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 1 << 20          # bytes either side; a multiple of 256 so that the payload keeps the allocator's alignment
CANARY = 0xA5


@contextlib.contextmanager
def guarded_allocations():
    real_empty, real_empty_like = torch.empty, torch.empty_like
    live = []

    def alloc(shape, dtype, device):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        raw = real_empty(GUARD + nbytes + pad + GUARD, dtype=torch.uint8, device=device)
        raw.fill_(CANARY)
        live.append((raw, nbytes))
        return raw[GUARD:GUARD + nbytes].view(dtype).view(*shape)

    def empty(*size, **kw):
        dev = kw.get("device")
        if dev is None or torch.device(dev).type != "cuda" or kw.get("dtype") not in (torch.uint8, torch.float32) or \
                any(k not in ("dtype", "device") for k in kw):
            return real_empty(*size, **kw)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        if len(shape) == 0:
            return real_empty(*size, **kw)
        return alloc(shape, kw["dtype"], dev)

    def empty_like(t, **kw):
        if kw or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            return real_empty_like(t, **kw)
        return alloc(tuple(t.shape), t.dtype, t.device)

    torch.empty, torch.empty_like = empty, empty_like
    try:
        yield live
    finally:
        torch.empty, torch.empty_like = real_empty, real_empty_like


def _check(live):
    torch.cuda.synchronize()
    assert len(live) >= 3, "the guarded allocator saw no library buffers"
    for raw, nbytes in live:
        lo, hi = raw[:GUARD], raw[GUARD + nbytes + ((-nbytes) % 256):]
        assert bool((lo == CANARY).all()), f"write below a {nbytes}-byte buffer"
        assert bool((hi == CANARY).all()), f"write above a {nbytes}-byte buffer"


@pytest.mark.parametrize("conv_mode", [0, 1, 2])
@pytest.mark.parametrize("family", ["int", "hgru", "int_noinh", "int_tanh"])
def test_int_family_writes_stay_inside_caller_buffers(family, conv_mode):
    from pathtracker_models_b200 import InT, ffhgru_hierarchy, synth_pathtracker
    torch.manual_seed(0)
    ctor, kw = {"int": (InT.InT, {}), "hgru": (ffhgru_hierarchy.FFhGRU, {}), "int_noinh": (InT.InT, {"no_inh": True}),
                "int_tanh": (InT.InT, {"nl": torch.tanh})}[family]
    m = ctor(32, 5, 7, **kw).cuda()
    m.conv_mode = conv_mode
    for (B, H, W, tm) in ((3, 32, 32, False), (2, 32, 64, True), (1, 64, 64, False)):
        x, y = synth_pathtracker(B, 5, H, W)
        x.requires_grad_(True)
        ref = m(x, testmode=True)[0] if tm else m(x)[0]
        ref.sum().backward()
        gref = [p.grad.clone() for p in m.parameters() if p.grad is not None]
        m.zero_grad()
        with guarded_allocations() as live:
            out = m(x, testmode=True)[0] if tm else m(x)[0]
            out.sum().backward()
            _check(live)
        # canary-filled buffers instead of whatever the allocator returned: no result may depend on a buffer's initial content
        assert torch.equal(out, ref)
        assert all(torch.equal(a, b) for a, b in zip(gref, [p.grad for p in m.parameters() if p.grad is not None]))
        m.zero_grad()


@pytest.mark.parametrize("conv_mode", [0, 1])
@pytest.mark.parametrize("filt,method,jp", [(15, "bptt", False), (7, "bptt", True), (15, "rbp", True), (7, "rbp", False)])
def test_convlstm_writes_stay_inside_caller_buffers(filt, method, jp, conv_mode):
    from pathtracker_models_b200 import convlstm
    from tests.convlstm_util import load_convlstm_case
    gabor = load_convlstm_case("convlstm_k7_small")["gabor"]
    torch.manual_seed(1)
    m = convlstm.ConvLSTM(timesteps=4, filt_size=filt, gabor=gabor, grad_method=method, num_iter=3, jacobian_penalty=jp).cuda().train()
    m.conv_mode = conv_mode
    for (B, H, W) in ((3, 32, 32), (1, 16, 64)):
        x = torch.rand(B, 1, H, W, device="cuda", requires_grad=True)
        tgt = (torch.rand(B, H, W, device="cuda") > 0.5).long()
        with guarded_allocations() as live:
            out, pen, loss = m(x, 0, 0, tgt, torch.nn.CrossEntropyLoss())
            (loss + (pen.mean() if jp else 0.0)).backward()
            _check(live)
        m.zero_grad()


@pytest.mark.parametrize("conv_mode", [0, 1])
def test_gru_writes_stay_inside_caller_buffers(conv_mode):
    from pathtracker_models_b200 import kys, synth_pathtracker
    torch.manual_seed(2)
    m = kys.GRU(dimensions=64, timesteps=3, kernel_size=7).cuda()
    m.conv_mode = conv_mode
    for (B, H, W) in ((2, 32, 32), (1, 16, 64)):
        x, y = synth_pathtracker(B, 3, H, W)
        x.requires_grad_(True)
        with guarded_allocations() as live:
            out = m(x)[0]
            out.sum().backward()
            _check(live)
        m.zero_grad()


@pytest.mark.parametrize("family", ["int", "hgru", "convlstm", "gru"])
def test_perturbed_cta_scheduling_changes_no_bits(family):
    """stand-in for racecheck / synccheck: the same forward + backward with and without a second stream saturating the SMs
    (large GEMMs), which changes which CTAs of a launch are co-resident, in what order they start and how far the
    programmatic-dependent-launch prologues run ahead.  The last-CTA folds, per-CTA accumulator slots and PDL waits must make
    the result independent of all that: every output and gradient bit-identical."""
    from pathtracker_models_b200 import InT, ffhgru_hierarchy, kys, convlstm, synth_pathtracker
    from tests.convlstm_util import load_convlstm_case
    torch.manual_seed(3)
    if family == "convlstm":
        m = convlstm.ConvLSTM(timesteps=6, filt_size=15, gabor=load_convlstm_case("convlstm_k7_small")["gabor"],
                              jacobian_penalty=True).cuda().train()
        x = torch.rand(8, 1, 32, 32, device="cuda")
        tgt = (torch.rand(8, 32, 32, device="cuda") > 0.5).long()

        def run():
            m.zero_grad()
            out, pen, loss = m(x, 0, 0, tgt, torch.nn.CrossEntropyLoss())
            (loss + pen.mean()).backward()
            return [out.detach().clone()] + [p.grad.clone() for p in m.parameters()]
    else:
        m = {"int": lambda: InT.InT(32, 8, 7), "hgru": lambda: ffhgru_hierarchy.FFhGRU(32, 8, 7),
             "gru": lambda: kys.GRU(dimensions=64, timesteps=8, kernel_size=7)}[family]().cuda()
        x, y = synth_pathtracker(16, 8, 32, 32)

        def run():
            m.zero_grad()
            out = m(x)[0]
            torch.nn.functional.binary_cross_entropy_with_logits(out, y.view_as(out)).backward()
            return [out.detach().clone()] + [p.grad.clone() for p in m.parameters() if p.grad is not None]
    quiet = run()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    a = torch.randn(6144, 6144, device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    for rep in range(3):
        with torch.cuda.stream(side):
            for _ in range(40 + 20 * rep):
                torch.mm(a, a, out=b)
        noisy = run()
        torch.cuda.synchronize()
        assert len(noisy) == len(quiet)
        for i, (u, v) in enumerate(zip(quiet, noisy)):
            assert torch.equal(u, v), f"tensor {i} differs under perturbed scheduling (repeat {rep})"
